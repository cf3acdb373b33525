"""Install the UNMODIFIED reference package into baseline/_ref (git-ignored; it travels to the GPU box with the snapshot).

    python oracle/install_reference.py            # needs /root/reference (the build container)

The reference is a bare script tree without setup.py / pyproject.toml, so the documented
``pip install --no-index --no-build-isolation --target baseline/_ref /root/reference`` fails ("Neither 'setup.py' nor
'pyproject.toml' found").  As the base contract allows for a read-only source tree, the install runs from a copy under /tmp to
which ONLY packaging metadata is added (a 4-line setup.py naming the package and its sub-packages); no reference source is
edited and none is copied into the repository's history.  ``bench.py --impl reference`` and the CPU-baseline leg import
``msa_tts.models.tacotron2nv`` / ``...tacotron2nv_loss`` / ``utils.grad_utils`` from there when it exists
(``cpu_baseline.kind = "reference"``); the meta-learning loop around them stays the restatement of oracle/meta.py because the
reference's own loop lives in ``higher`` (absent, no network).  TEST / BENCH INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")

SETUP = '''from setuptools import setup, find_packages
setup(name="msa_tts_reference", version="0.0.0", packages=find_packages(include=["msa_tts", "msa_tts.*"]),
      description="HamedHemati/MetaSpeakerAdaptation-TTS, unmodified sources (packaging metadata only)")
'''


def install(force: bool = False) -> str:
    marker = os.path.join(DST, "msa_tts", "models", "tacotron2nv.py")
    if os.path.exists(marker) and not force:
        return DST
    if not os.path.isdir(SRC):
        raise RuntimeError(f"{SRC} is not present: the reference can only be installed in the build container")
    tmp = tempfile.mkdtemp(prefix="msa_ref_src_")
    try:
        work = os.path.join(tmp, "reference")
        shutil.copytree(SRC, work)
        with open(os.path.join(work, "setup.py"), "w") as f:
            f.write(SETUP)
        os.makedirs(DST, exist_ok=True)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps", "--upgrade",
               "--find-links", "/opt/wheelhouse", "--target", DST, work]
        subprocess.check_call(cmd)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if not os.path.exists(marker):
        raise RuntimeError("reference install did not produce msa_tts/models/tacotron2nv.py")
    return DST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
