"""ctypes binding of the in-tree C-ABI library (include/mirror_b200.h).

The product path has NO fallback: if the library cannot be loaded (or built
from the in-tree sources with nvcc) importing any kernel raises immediately.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_VARIANT = os.environ.get("MIRROR_B200_VARIANT", "")  # A/B builds for kernel experiments (see build.py)
LIB_PATH = os.path.join(_HERE, f"libmirror_b200{'_' + _VARIANT if _VARIANT else ''}.so")
_lib = None


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("a", ctypes.c_void_p), ("b", ctypes.c_void_p),
        ("a_mn_major", ctypes.c_int32), ("b_mn_major", ctypes.c_int32),
        ("lda", ctypes.c_int64), ("ldb", ctypes.c_int64),
        ("a_bs1", ctypes.c_int64), ("a_bs2", ctypes.c_int64), ("b_bs1", ctypes.c_int64), ("b_bs2", ctypes.c_int64),
        ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32),
        ("batch1", ctypes.c_int32), ("batch2", ctypes.c_int32),
        ("alpha", ctypes.c_float),
        ("bias", ctypes.c_void_p),
        ("act", ctypes.c_int32),
        ("drop_p", ctypes.c_float),
        ("drop_seed", ctypes.c_uint64),
        ("res", ctypes.c_void_p),
        ("res_is_bf16", ctypes.c_int32),
        ("gamma", ctypes.c_float),
        ("ldr", ctypes.c_int64), ("r_bs1", ctypes.c_int64), ("r_bs2", ctypes.c_int64),
        ("beta", ctypes.c_float),
        ("out_f32", ctypes.c_void_p),
        ("ldc32", ctypes.c_int64), ("c32_bs1", ctypes.c_int64), ("c32_bs2", ctypes.c_int64),
        ("out_bf16", ctypes.c_void_p),
        ("ldc16", ctypes.c_int64), ("c16_bs1", ctypes.c_int64), ("c16_bs2", ctypes.c_int64),
        ("split_k", ctypes.c_int32),
        ("diag", ctypes.c_float),
        ("res2", ctypes.c_void_p),
        ("gamma2", ctypes.c_float),
        ("res_row_div", ctypes.c_int32),
        ("mode", ctypes.c_int32),
        ("stats", ctypes.c_void_p),
    ]


class FlashArgs(ctypes.Structure):
    _fields_ = [
        ("x", ctypes.c_void_p), ("y", ctypes.c_void_p), ("v", ctypes.c_void_p),
        ("x_ld", ctypes.c_int64), ("x_hs", ctypes.c_int64), ("x_bs", ctypes.c_int64),
        ("y_ld", ctypes.c_int64), ("y_hs", ctypes.c_int64), ("y_bs", ctypes.c_int64),
        ("v_ld", ctypes.c_int64), ("v_hs", ctypes.c_int64), ("v_bs", ctypes.c_int64),
        ("R", ctypes.c_int32), ("C", ctypes.c_int32), ("d", ctypes.c_int32), ("heads", ctypes.c_int32), ("batch", ctypes.c_int32),
        ("alpha", ctypes.c_float),
        ("out", ctypes.c_void_p), ("o_ld", ctypes.c_int64), ("o_hs", ctypes.c_int64), ("o_bs", ctypes.c_int64),
        ("res", ctypes.c_void_p), ("r_ld", ctypes.c_int64), ("r_hs", ctypes.c_int64), ("r_bs", ctypes.c_int64),
        ("lse2", ctypes.c_void_p),
    ]


class FlashOut(ctypes.Structure):
    _fields_ = [
        ("ptr", ctypes.c_void_p), ("is_f32", ctypes.c_int32),
        ("ld", ctypes.c_int64), ("hs", ctypes.c_int64), ("bs", ctypes.c_int64),
        ("res", ctypes.c_void_p), ("r_ld", ctypes.c_int64), ("r_hs", ctypes.c_int64), ("r_bs", ctypes.c_int64),
        ("row_div", ctypes.c_int32), ("rscale", ctypes.c_float),
    ]


class FlashBwdArgs(ctypes.Structure):
    _fields_ = [
        ("a", ctypes.c_void_p), ("b", ctypes.c_void_p), ("c", ctypes.c_void_p), ("dd", ctypes.c_void_p),
        ("a_ld", ctypes.c_int64), ("a_hs", ctypes.c_int64), ("a_bs", ctypes.c_int64),
        ("b_ld", ctypes.c_int64), ("b_hs", ctypes.c_int64), ("b_bs", ctypes.c_int64),
        ("c_ld", ctypes.c_int64), ("c_hs", ctypes.c_int64), ("c_bs", ctypes.c_int64),
        ("d_ld", ctypes.c_int64), ("d_hs", ctypes.c_int64), ("d_bs", ctypes.c_int64),
        ("T", ctypes.c_int32), ("L", ctypes.c_int32), ("d", ctypes.c_int32), ("heads", ctypes.c_int32), ("batch", ctypes.c_int32),
        ("cols", ctypes.c_int32),
        ("alpha", ctypes.c_float),
        ("lse2", ctypes.c_void_p), ("dot", ctypes.c_void_p),
        ("out1", FlashOut), ("out2", FlashOut),
    ]


def lib():
    """Load the library, (re)building it first when it is missing or older than its sources.  The digest check is cheap
    (hash of csrc/ + the header) and runs under an exclusive file lock: the ranks of a DDP job start together and must not
    compile into the same object directory at once, and an edited .cu must never run as a stale .so."""
    global _lib
    if _lib is None:
        from . import build as _build  # compile in-tree; raises if nvcc is missing
        import fcntl
        with open(os.path.join(_HERE, ".build.lock"), "w") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if os.environ.get("MIRROR_B200_NO_REBUILD") != "1" or not os.path.exists(LIB_PATH):
                    _build.build()
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.mirror_last_error.restype = ctypes.c_char_p
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"mirror_b200 {what} failed (code {rc}): {lib().mirror_last_error().decode()}")


# ---- argument tables for the non-GEMM entry points (see include/mirror_b200.h) ----
_P, _I32, _I64, _F, _U64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_uint64
SIGNATURES = {
    "mirror_gemm_bf16": [_P, _P],
    "mirror_gemm_bf16_multi": [_P, _I32, _P],
    "mirror_gemm_nparts": [_I32],
    "mirror_gemm_bf16_simt": [_P, _P],
    "mirror_cast_f32_bf16": [_P, _I64, _I32, _I64, _P, _I32, _I64, _P],
    "mirror_cast_split3": [_P, _I64, _I32, _I64, _P, _I64, _I32, _I32, _I32, _P],
    "mirror_copy_rows_f32": [_P, _I64, _I64, _I32, _P, _I64, _P],
    "mirror_axpy_f32": [_P, _P, _I64, _F, _P],
    "mirror_act_fwd": [_P, _I64, _I32, _F, _U64, _P, _P, _P],
    "mirror_act_bwd": [_P, _I64, _I64, _P, _I64, _I64, _I32, _I32, _I32, _I32, _F, _U64, _P, _I64, _I64, _P, _I64, _I64, _P],
    "mirror_wsi_assemble_fwd": [_P, _P, _I32, _I32, _I32, _I32, _P],
    "mirror_wsi_embed_bwd": [_P, _P, _I32, _I32, _I32, _I32, _P, _P, _P],
    "mirror_rank_mask": [_P, _I32, _I32, _I32, _P, _P],
    "mirror_mask_pos_fwd": [_P, _P, _P, _I32, _P, _I32, _I32, _I32, _I32, _P],
    "mirror_mask_pos_bwd": [_P, _P, _P, _P, _I32, _P, _I32, _I32, _I32, _I32, _P],
    "mirror_landmark_fwd": [_P, _P, _I32, _I32, _I32, _I32, _I32, _P],
    "mirror_colsum": [_P, _I32, _I64, _I32, _I64, _P, _P],
    "mirror_reparam_fwd": [_P, _P, _P, _I64, _P, _P, _P],
    "mirror_reparam_bwd": [_P, _P, _P, _I64, _P, _P, _P],
    "mirror_pinv_init_softmax_bwd": [_P, _P, _P, _P, _I32, _I32, _P, _F, _P, _P],
    "mirror_rowdot_bf16": [_P, _P, _P, _I64, _I32, _P, _P],
    "mirror_token_fanout_bwd": [_P, _P, _P, _I64, _I64, _I32, _I32, _I32, _P, _P],
    "mirror_layernorm_fwd": [_P, _P, _P, _F, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P],
    "mirror_layernorm_bwd": [_P, _I32, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P],
    "mirror_softmax_fwd": [_P, _I64, _I32, _P, _P, _P],
    "mirror_softmax_bwd": [_P, _P, _I64, _I32, _F, _P, _P, _P],
    "mirror_l2norm_fwd": [_P, _I64, _I32, _I32, _F, _P, _P, _I64, _P, _P],
    "mirror_l2norm_bwd": [_P, _I64, _P, _I64, _P, _I32, _I32, _P, _I64, _I32, _P],
    "mirror_res_conv_fwd": [_P, _P, _I32, _I32, _I32, _P, _P],
    "mirror_res_conv_bwd": [_P, _P, _P, _I32, _I32, _I32, _P, _P, _P],
    "mirror_pinv_init": [_P, _I32, _I32, _P, _P, _P, _P],
    "mirror_pinv_init_bwd": [_P, _P, _I32, _I32, _P, _P, _I32, _P],
    "mirror_ppeg_fwd": [_P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _P, _P, _P, _P],
    "mirror_ppeg_bwd": [_P, _P, _P, _I32, _I32, _I32, _P, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "mirror_rna_attn_fwd": [_P, _I32, _I32, _P, _P, _P],
    "mirror_rna_attn_bwd": [_P, _P, _I32, _I32, _P, _P, _P],
    "mirror_flash_softmax_pv": [_P, _P],
    "mirror_flash_bwd": [_P, _P],
    "mirror_debug_flash_trace": [_P, _I64],
    "mirror_contrastive_nsplit": [_I32, _I32],
    "mirror_contrastive_stats": [_P, _I64, _P, _I64, _I32, _I32, _I32, _P, _I32, _P, _I32, _P, _P, _P],
    "mirror_contrastive_grad": [_P, _I64, _P, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P, _I32, _P, _P, _P, _P, _P, _I64, _P, _P],
    "mirror_contrastive_loss": [_P, _P, _P, _I32, _F, _F, _F, _P, _P, _P],
    "mirror_contrastive_coef": [_P, _I32, _I32, _F, _F, _F, _P, _P, _P],
    "mirror_masked_mse_fwd": [_P, _I64, _P, _I64, _P, _I32, _I32, _I32, _P, _P, _P],
    "mirror_masked_mse_bwd": [_P, _I64, _P, _I64, _P, _I32, _I32, _I32, _P, _P, _F, _P, _I64, _I32, _P, _I64, _I32, _P],
    "mirror_gauss_kl_fwd": [_P, _P, _I64, _I32, _P, _P],
    "mirror_gauss_kl_bwd": [_P, _P, _I64, _I32, _P, _F, _P, _P, _P],
    "mirror_sym_kl_fwd": [_P, _I32, _I32, _P, _P],
    "mirror_sym_kl_bwd": [_P, _I32, _I32, _P, _F, _P, _P, _P],
    "mirror_loss_combine": [_P, _P, _P, _P],
    "mirror_set_dropout_epoch": [_P],
    "mirror_gather_rows": [_P, _I32, _I64, _I64, _P, _I64, _I32, _P, _P],
    "mirror_adam_step": [_P, _P, _P, _P, _I64, _P, _F, _F, _F, _F, _I32, _P, _P, _P],
    "mirror_grad_sumsq": [_P, _I64, _P, _P],
    "mirror_tail_scalars": [_P, _F, _P, _P, _P, _F, _F, _P],
}
EXPORTS = ["mirror_last_error", "mirror_abi_version", "mirror_device_supported"] + list(SIGNATURES)


def fn(name):
    f = getattr(lib(), name)
    if f.argtypes is None:
        f.argtypes = SIGNATURES[name]
        f.restype = ctypes.c_int
    return f
