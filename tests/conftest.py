import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # a "gpu" test on a box without CUDA is an error in the harness, not a skip: fail loudly
    pass


GOLDEN = os.path.join(ROOT, "tests", "golden")
