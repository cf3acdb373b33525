"""ctypes binding of libmsa_b200.so (include/msa_b200.h).  Fails loudly: no fallback of any kind."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MSA_LIB_PATH: development only (A/B runs of differently compiled builds of the same library)
LIB_PATH = os.environ.get("MSA_LIB_PATH") or os.path.join(HERE, "libmsa_b200.so")

c_f32p = C.c_void_p
c_i64p = C.c_void_p
c_u8p = C.c_void_p


class MsaConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_symbols", "enc_dim", "enc_kernel", "enc_n_convs", "spk_mode", "spk_in_dim", "spk_dim", "num_speakers",
        "n_mel", "prenet_dim", "attn_rnn_dim", "dec_rnn_dim", "attn_dim", "loc_filters", "loc_kernel", "post_dim",
        "post_kernel", "post_n_convs", "attn_norm", "forward_attn", "trans_agent", "windowing", "forward_attn_mask",
        "max_decoder_steps", "early_stopping", "loss_reduction", "gemm_tf32")] + [(n, C.c_float) for n in (
        "p_attn_dropout", "p_dec_dropout", "gate_threshold", "loss_pos_weight")] + [(n, C.c_int32) for n in (
        "freeze_charemb", "freeze_encoder", "freeze_decoder", "residual_encoder")]


# name -> (restype, argtypes); every symbol declared in include/msa_b200.h
V, I, I64, F, SZ, U64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t, C.c_uint64
SIGNATURES = {
    "msa_last_error_string": (C.c_char_p, []),
    "msa_version": (I, []),
    "msa_create": (I, [C.POINTER(MsaConfig), I, C.POINTER(V)]),
    "msa_destroy": (I, [V]),
    "msa_sm_count": (I, [V]),
    "msa_launch_count": (C.c_longlong, []),
    "msa_profile_enable": (I, [V, I]),
    "msa_profile_kernels": (I, []),
    "msa_profile_name": (C.c_char_p, [I]),
    "msa_profile_read": (I, [V, C.POINTER(C.c_double), C.POINTER(I64)]),
    "msa_param_count": (I, [V]),
    "msa_param_total": (I64, [V]),
    "msa_param_info": (I, [V, I, C.POINTER(C.c_char_p), C.POINTER(I64), C.POINTER(I64)]),
    "msa_bn_count": (I, [V]),
    "msa_bn_total": (I64, [V]),
    "msa_bn_info": (I, [V, I, C.POINTER(I64), C.POINTER(C.c_int32)]),
    "msa_mask_count": (I, [V]),
    "msa_mask_total": (I64, [V, I, I, I]),
    "msa_mask_info": (I, [V, I, I, I, I, C.POINTER(C.c_char_p), C.POINTER(I64), C.POINTER(I64), C.POINTER(F)]),
    "msa_masks_generate": (I, [V, V, I, I, I, U64, V]),
    "msa_workspace_bytes": (SZ, [V, I, I, I]),
    "msa_train_forward": (I, [V, V, SZ, V, V, V, V, V, V, V, V, V, V, I, I, I, V, V, V, V, V, V]),
    "msa_train_backward": (I, [V, V, SZ, V, V, V, V, V, I, F, V]),
    "msa_group_workspace_bytes": (SZ, [V, I, I, I, I]),
    "msa_train_forward_group": (I, [V, I, V, SZ, V, V, V, V, V, V, V, V, V, V, I, I, I, V, V]),
    "msa_train_backward_group": (I, [V, V, SZ, V, V, I, F, V]),
    "msa_backward_mark_event": (I, [V, V, C.POINTER(I64)]),
    "msa_train_loss": (I, [V, V, V, V, I, F, V, V]),
    "msa_train_mcd": (I, [V, V, V, I, V, V]),
    "msa_loss_scratch_floats": (SZ, [I, I, I]),
    "msa_tacotron2_loss": (I, [V, V, V, V, V, V, I, I, I, I, F, V, V, V, V, V, V]),
    "msa_loss_grads": (I, [V, V, V, V, V, V]),
    "msa_check_abort": (I, [V, V, V]),
    "msa_abort_guard": (I, [V, V, V]),
    "msa_abort_read_async": (I, [V, V, V]),
    "msa_abort_clear": (I, [V, V]),
    "msa_debug_raise_abort": (I, [V, V]),
    "msa_profile_phases": (I, [V, V, I, C.POINTER(I64), I]),
    "msa_profile_trace_step": (I, [V, I]),
    "msa_profile_trace": (I, [V, V, I, C.POINTER(I64), I]),
    "msa_get_buffer": (I, [V, V, C.c_char_p, C.POINTER(V), C.POINTER(I64)]),
    "msa_infer_workspace_bytes": (SZ, [V, I, I, I]),
    "msa_infer": (I, [V, V, SZ, V, V, V, V, V, V, V, I, I, I, V, V, V, V, V]),
    "msa_gemm_nt_scratch_floats": (SZ, [I64, I64, I64]),
    "msa_gemm_nt": (I, [I64, I64, I64, F, V, I64, V, I64, F, V, I64, I, V, V]),
    "msa_gemm": (I, [I, I, I64, I64, I64, F, V, I64, V, I64, F, V, I64, I, V, V]),
    "msa_conv1d_scratch_floats": (SZ, [I, I, I, I, I]),
    "msa_conv1d_stat_slabs": (I, [I, I, I, I, I, I]),
    "msa_conv1d_repack": (I, [V, V, I, I, I, V]),
    "msa_conv1d_fwd": (I, [V, I, I, I, V, I, I, V, V, I, V, V, V]),
    "msa_conv1d_dx": (I, [V, I, I, I, V, I, I, V, I, V, V]),
    "msa_conv1d_dw": (I, [V, V, I, I, I, I, I, F, I, V, I, V]),
    "msa_flat_sgd_step": (I, [V, V, V, V, I64, F, F, F, F, I, I, V]),
    "msa_flat_axpy": (I, [V, V, I64, F, I, V]),
    "msa_flat_reptile_delta": (I, [V, V, V, I64, F, I, V]),
    "msa_flat_partials": (I, []),
    "msa_flat_sumsq": (I, [V, I64, V, V, V]),
    "msa_flat_clip_sgd": (I, [V, V, V, V, I64, F, F, F, F, F, I, I, V]),
    "msa_flat_clip_adam": (I, [V, V, V, V, V, I64, F, F, F, F, F, I, F, V]),
    "msa_flat_adam_step": (I, [V, V, V, V, V, I64, F, F, F, F, F, I, V]),
    "msa_ewc_fisher_accum": (I, [V, V, I64, F, I, V]),
    "msa_ewc_penalty": (I, [V, V, V, I64, V, V, V]),
    "msa_ewc_sgd_step": (I, [V, V, V, V, I64, F, F, V, V, V]),
    "msa_ewc_penalty_grad": (I, [V, V, V, V, I64, F, V, V, V]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build`; "
                               "this package has no CPU / PyTorch fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the ABI and the header drift apart
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().msa_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"libmsa_b200 {what} failed with code {rc}: {msg}")
