"""ctypes binding of libb200timm.so (the C ABI declared in include/b200timm.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C csrc``.  There is no fallback: if
the shared object is missing, loading raises; if a kernel reports an error, ``B200Error`` is raised.
"""
import ctypes
import os
import re
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200TIMM_LIB: load another build of the library instead (same-box A/B runs of two kernel versions)
LIB_PATH = os.environ.get("B200TIMM_LIB") or os.path.join(_HERE, "libb200timm.so")
HEADER_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "include", "b200timm.h"))

OK, ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_CUDA = 0, 1, 2, 3
DTYPE_BF16, DTYPE_F32, DTYPE_U8 = 0, 1, 2
EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_DGELU, EPI_BIAS_ROPE = 0, 1, 2, 3, 4


class B200Error(RuntimeError):
    """A libb200timm entry point returned a non-zero code."""


class GemmDesc(ctypes.Structure):
    """Mirror of ``struct b200_gemm_desc``."""

    _fields_ = [
        ("a", ctypes.c_void_p),
        ("b", ctypes.c_void_p),
        ("out", ctypes.c_void_p),
        ("bias", ctypes.c_void_p),
        ("aux", ctypes.c_void_p),
        ("gamma", ctypes.c_void_p),
        ("out2", ctypes.c_void_p),
        ("row_scale", ctypes.c_void_p),
        ("lda", ctypes.c_int64),
        ("ldb", ctypes.c_int64),
        ("ldo", ctypes.c_int64),
        ("ldaux", ctypes.c_int64),
        ("ldo2", ctypes.c_int64),
        ("M", ctypes.c_int32),
        ("N", ctypes.c_int32),
        ("K", ctypes.c_int32),
        ("a_mn", ctypes.c_int32),
        ("b_mn", ctypes.c_int32),
        ("epilogue", ctypes.c_int32),
        ("out_dtype", ctypes.c_int32),
        ("bias_dtype", ctypes.c_int32),
        ("gamma_dtype", ctypes.c_int32),
        ("accumulate", ctypes.c_int32),
        ("rows_per_scale", ctypes.c_int32),
        ("split_k", ctypes.c_int32),
        ("rope", ctypes.c_void_p),
        ("rope_tokens", ctypes.c_int32),
        ("rope_prefix", ctypes.c_int32),
        ("rope_cols", ctypes.c_int32),
        ("act", ctypes.c_int32),
        ("colsum", ctypes.c_void_p),
        ("ln_stats_out", ctypes.c_void_p),
        ("ln_stats", ctypes.c_void_p),
        ("ln_c1", ctypes.c_void_p),
        ("ln_slots", ctypes.c_int32),
        ("ln_eps", ctypes.c_float),
    ]


class OptimTensor(ctypes.Structure):
    """Mirror of ``struct b200_optim_tensor``."""

    _fields_ = [("param", ctypes.c_void_p), ("grad", ctypes.c_void_p), ("state1", ctypes.c_void_p),
                ("state2", ctypes.c_void_p), ("numel", ctypes.c_int64)]


_vp, _i32, _i64, _f32, _f64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_double
_otp = ctypes.POINTER(OptimTensor)

# name -> (restype, argtypes); must cover every B200_API prototype of include/b200timm.h
# (tests/test_cabi.py checks the two lists against each other and against the .so's exports).
SIGNATURES = {
    "b200_last_error": (ctypes.c_char_p, []),
    "b200_abi_version": (ctypes.c_int, []),
    "b200_device_check": (ctypes.c_int, []),
    "b200_gemm": (ctypes.c_int, [ctypes.POINTER(GemmDesc), _vp]),
    "b200_gemm_ln_slots": (ctypes.c_int32, [_i32, _i32]),
    "b200_set_trace": (None, [_i32, _vp]),
    "b200_rope_qk": (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b200_layernorm_fwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _i32, _i32, _i32, _vp]),
    "b200_layernorm_bwd_ws_rows": (ctypes.c_int, [_i64, _i32]),
    "b200_layernorm_bwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp]),
    "b200_rmsnorm_fwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _f32, _i32, _i32, _i32, _vp]),
    "b200_rmsnorm_bwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp]),
    "b200_convert_f32": (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32, _vp]),
    "b200_attn_fwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp]),
    "b200_attn_bwd_ws_floats": (ctypes.c_int64, [_i32, _i32, _i32]),
    "b200_attn_bwd_ws_floats_hd": (ctypes.c_int64, [_i32, _i32, _i32, _i32]),
    "b200_attn_bwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp, _vp]),
    "b200_attn_fwd_bias": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _f32, _vp]),
    "b200_attn_bwd_bias": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _f32,
                                          _vp, _vp]),
    "b200_patchify": (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b200_patchify_u8": (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, ctypes.POINTER(ctypes.c_float),
                                        ctypes.POINTER(ctypes.c_float), _vp]),
    "b200_assemble_tokens": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b200_naflex_assemble": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b200_naflex_assemble_bwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "b200_scale_rows_cols": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "b200_colsum_mul": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "b200_colsum": (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32, _vp]),
    "b200_dwconv_nhwc": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b200_dwconv_nhwc_ln": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b200_dwconv_nhwc_wgrad": (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b200_unpatchify_nhwc": (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b200_patchify_nhwc": (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b200_avgpool_nhwc": (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _vp]),
    "b200_qk_norm_fwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _f32, _i32, _vp]),
    "b200_qk_norm_bwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _i32, _i32, _i32, _vp]),
    "b200_grn_fwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp]),
    "b200_grn_bwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _vp]),
    "b200_glu_fwd": (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "b200_glu_bwd": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "b200_sgd_step": (ctypes.c_int, [_otp, _i32, _i32, _f64, _f64, _f64, _f64, _i32, _i32, _vp, _vp]),
    "b200_adam_step": (ctypes.c_int, [_otp, _i32, _i32, _f64, _f64, _f64, _f64, _f64, _i32, _i32, _vp, _vp]),
    "b200_grad_sumsq": (ctypes.c_int, [_otp, _i32, _i32, _vp, _vp]),
    "b200_clip_coef": (ctypes.c_int, [_vp, _f64, _vp, _vp, _vp]),
    "b200_grad_scale": (ctypes.c_int, [_otp, _i32, _i32, _vp, _vp]),
    "b200_ema_update": (ctypes.c_int, [_otp, _i32, _i32, _i32, _f64, _vp]),
    "b200_patch_embed_fwd": (ctypes.c_int, [_vp, _vp, _vp, _i32, _vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b200_batch_patchify": (ctypes.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "b200_patches_u8": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "b200_act_bwd": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int64, _i32, _vp]),
    "b200_grad_publish": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, _vp]),
    "b200_grad_push": (ctypes.c_int, [_vp, _i32, _vp, ctypes.c_int64, _f32, _vp]),
    "b200_cross_entropy_fwd": (ctypes.c_int, [_vp, _i32, _vp, _vp, _i32, _f64, ctypes.c_int64, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    "b200_cross_entropy_bwd": (ctypes.c_int, [_vp, _i32, _vp, _vp, _i32, _f64, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
}

_lib = None
_lock = threading.Lock()


def header_symbols():
    """Names of all B200_API functions declared in include/b200timm.h."""
    with open(HEADER_PATH) as f:
        text = f.read()
    return sorted(set(re.findall(r"B200_API[^;(]*?\b(b200_\w+)\s*\(", text)))


def load():
    """Load libb200timm.so (once) and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise B200Error(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(there is no CPU / PyTorch fallback for the B200 hot path)")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc, what):
    if rc != OK:
        msg = load().b200_last_error()
        raise B200Error(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
