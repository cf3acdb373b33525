"""The C-ABI library loads and exports every symbol include/msa_b200.h declares (no compute calls: runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "msa_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msa_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_documented_groups():
    names = _declared()
    for must in ("msa_create", "msa_destroy", "msa_train_forward", "msa_train_backward", "msa_infer", "msa_flat_sgd_step",
                 "msa_flat_axpy", "msa_flat_reptile_delta", "msa_flat_clip_adam", "msa_flat_adam_step", "msa_ewc_sgd_step", "msa_tacotron2_loss",
                 "msa_train_mcd", "msa_gemm_nt", "msa_masks_generate",
                 "msa_last_error_string"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from msa_tts_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        pytest.fail(f"{_lib.LIB_PATH} missing: run `python __graft_entry__.py build` first")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    unbound = [n for n in _declared() if n not in _lib.SIGNATURES]
    assert not unbound, f"declared in the header but not bound in _lib.py: {unbound}"


def test_no_device_is_a_loud_error_not_a_fallback():
    """Without a CUDA device msa_create fails with MSA_E_NODEVICE and says so (there is no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU box: covered by the gpu tests")
    import msa_tts_b200 as pkg
    from msa_tts_b200 import _lib
    from msa_tts_b200.engine import Engine, make_c_config
    lib = _lib.load()
    cfg = make_c_config(pkg.small_params())
    h = ctypes.c_void_p()
    rc = lib.msa_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc == -4 and b"no CPU fallback" in lib.msa_last_error_string()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(pkg.small_params())
