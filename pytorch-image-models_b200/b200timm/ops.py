"""Tensor-level wrappers over the C ABI (raw forward / backward kernels, no autograd).

PyTorch is used here only for device memory (caching allocator) and the current CUDA stream; every
numerical operation is a libb200timm kernel.  All functions require CUDA tensors and raise on anything
the kernels do not support (no fallback).
"""
import ctypes

import torch

from . import _lib
from ._lib import (DTYPE_BF16, DTYPE_F32, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_DGELU, B200Error,
                   GemmDesc)

__all__ = [
    "layernorm_fwd", "layernorm_bwd", "gemm", "linear_fwd", "attention_fwd", "patchify", "assemble_tokens",
    "colsum", "EPI_BIAS", "EPI_BIAS_GELU", "EPI_BIAS_RESIDUAL", "EPI_DGELU",
]


def _stream():
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------------------------ launch accounting
_launches = 0          # kernels of libb200timm launched since reset_counters()
_gemm_timing = None    # None, or a list of (start_event, end_event, flops) while bench.py times the GEMMs


def _count(n=1):
    global _launches
    _launches += n


def reset_counters():
    global _launches
    _launches = 0


def launch_count():
    return _launches


def enable_gemm_timing(flag):
    """bench.py: bracket every b200_gemm launch with CUDA events on the launching stream (roofline.achieved)."""
    global _gemm_timing
    _gemm_timing = [] if flag else None


def gemm_timing_summary():
    """(total algorithmic FLOPs, total milliseconds, number of launches) of the GEMMs timed so far."""
    if not _gemm_timing:
        return 0.0, 0.0, 0
    torch.cuda.synchronize()
    flops = sum(f for _, _, f in _gemm_timing)
    ms = sum(s.elapsed_time(e) for s, e, _ in _gemm_timing)
    return flops, ms, len(_gemm_timing)


def _dt(t):
    if t.dtype == torch.bfloat16:
        return DTYPE_BF16
    if t.dtype == torch.float32:
        return DTYPE_F32
    raise B200Error(f"unsupported dtype {t.dtype} (bf16 / fp32 only)")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise B200Error("libb200timm kernels need CUDA tensors (there is no CPU path)")


def _param_dtype(*ps):
    ds = {p.dtype for p in ps if p is not None}
    if len(ds) > 1:
        raise B200Error(f"parameters must share one dtype, got {ds}")
    return _dt(next(p for p in ps if p is not None)) if ds else DTYPE_BF16


# --------------------------------------------------------------------------------------------- LayerNorm
def layernorm_fwd(x, weight, bias, eps, out_dtype=None, save_stats=False):
    """y = LayerNorm(x) over the last dim.  Returns (y, mean, rstd) (stats are None unless save_stats)."""
    _need_cuda(x, weight, bias)
    cols = x.shape[-1]
    x2 = x.reshape(-1, cols)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    rows = x2.shape[0]
    y = torch.empty(x2.shape, dtype=out_dtype or x.dtype, device=x.device)
    mean = rstd = None
    if save_stats:
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    rc = lib.b200_layernorm_fwd(x2.data_ptr(), _ptr(weight), _ptr(bias), y.data_ptr(), _ptr(mean), _ptr(rstd), rows,
                                cols, float(eps), _dt(x2), _dt(y), _param_dtype(weight, bias), _stream())
    _lib.check(rc, "b200_layernorm_fwd")
    _count()
    return y.view(*x.shape[:-1], cols), mean, rstd


def layernorm_bwd(dy, x, weight, mean, rstd, need_param_grads=True):
    """Returns (dx, dweight fp32, dbias fp32)."""
    _need_cuda(dy, x, weight, mean, rstd)
    cols = x.shape[-1]
    x2 = x.reshape(-1, cols)
    dy2 = dy.reshape(-1, cols)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    if not dy2.is_contiguous():
        dy2 = dy2.contiguous()
    rows = x2.shape[0]
    lib = _lib.load()
    ws_rows = lib.b200_layernorm_bwd_ws_rows(rows, cols)
    ws = torch.empty(ws_rows, 2 * cols, dtype=torch.float32, device=x.device)
    dx = torch.empty_like(x2)
    dw = db = None
    if need_param_grads:
        dw = torch.empty(cols, dtype=torch.float32, device=x.device)
        db = torch.empty(cols, dtype=torch.float32, device=x.device)
    rc = lib.b200_layernorm_bwd(dy2.data_ptr(), x2.data_ptr(), _ptr(weight), mean.data_ptr(), rstd.data_ptr(),
                                dx.data_ptr(), _ptr(dw), _ptr(db), ws.data_ptr(), rows, cols, _dt(x2), _dt(dy2),
                                _param_dtype(weight), _stream())
    _lib.check(rc, "b200_layernorm_bwd")
    _count(2 if need_param_grads else 1)
    return dx.view(x.shape), dw, db


# --------------------------------------------------------------------------------------------- GEMM
def _check_2d_bf16(t, name):
    if t.dim() != 2 or t.dtype != torch.bfloat16 or t.stride(1) != 1:
        raise B200Error(f"{name}: need a 2-D bf16 tensor with unit inner stride, got {tuple(t.shape)} {t.dtype} "
                        f"strides {t.stride()}")


def gemm(a, b, M, N, K, *, a_mn=False, b_mn=False, epilogue=EPI_BIAS, bias=None, aux=None, gamma=None, out=None,
         out_dtype=torch.bfloat16, accumulate=False, split_k=1):
    """out[M,N] = epilogue(sum_k A[m,k] B[n,k]).

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True) bf16; b: [N,K] (b_mn=False) or [K,N] (b_mn=True) bf16.
    """
    _need_cuda(a, b, bias, aux, gamma, out)
    _check_2d_bf16(a, "a")
    _check_2d_bf16(b, "b")
    exp_a = (K, M) if a_mn else (M, K)
    exp_b = (K, N) if b_mn else (N, K)
    if tuple(a.shape) != exp_a or tuple(b.shape) != exp_b:
        raise B200Error(f"gemm: operand shapes {tuple(a.shape)}, {tuple(b.shape)} do not match M={M} N={N} K={K} "
                        f"a_mn={a_mn} b_mn={b_mn}")
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=a.device)
    elif out.dim() != 2 or tuple(out.shape) != (M, N) or out.stride(1) != 1:
        raise B200Error("gemm: out must be [M,N] with unit inner stride")
    d = GemmDesc()
    d.a, d.b, d.out = a.data_ptr(), b.data_ptr(), out.data_ptr()
    d.bias, d.aux, d.gamma = _ptr(bias), _ptr(aux), _ptr(gamma)
    d.lda, d.ldb, d.ldo = a.stride(0), b.stride(0), out.stride(0)
    if aux is not None:
        _check_2d_bf16(aux, "aux")
        if tuple(aux.shape) != (M, N):
            raise B200Error("gemm: aux must be [M,N]")
        d.ldaux = aux.stride(0)
    d.M, d.N, d.K = M, N, K
    d.a_mn, d.b_mn = int(a_mn), int(b_mn)
    d.epilogue = epilogue
    d.out_dtype = _dt(out)
    d.bias_dtype = _dt(bias) if bias is not None else DTYPE_BF16
    d.gamma_dtype = _dt(gamma) if gamma is not None else DTYPE_BF16
    d.accumulate = int(accumulate)
    d.split_k = int(split_k)
    if _gemm_timing is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    rc = _lib.load().b200_gemm(ctypes.byref(d), _stream())
    _lib.check(rc, "b200_gemm")
    if _gemm_timing is not None:
        ev1.record()
        _gemm_timing.append((ev0, ev1, 2.0 * M * N * K))
    _count()
    return out


def linear_fwd(x2d, weight, bias=None, epilogue=EPI_BIAS, aux=None, gamma=None, out=None):
    """F.linear on a [M,K] bf16 activation with a fused epilogue (weight: [N,K] bf16)."""
    M, K = x2d.shape
    N = weight.shape[0]
    return gemm(x2d, weight, M, N, K, epilogue=epilogue, bias=bias, aux=aux, gamma=gamma, out=out)


# --------------------------------------------------------------------------------------------- attention
def attention_fwd(qkv, B, N, H, hd, scale, kv_len=None, need_lse=False):
    """qkv: bf16 [B*N, 3*H*hd] (or [B,N,3*H*hd]) contiguous.  Returns (out [B,N,H*hd] bf16, lse [B,H,N] | None)."""
    _need_cuda(qkv, kv_len)
    if qkv.dtype != torch.bfloat16 or not qkv.is_contiguous() or qkv.numel() != B * N * 3 * H * hd:
        raise B200Error("attention_fwd: qkv must be a contiguous bf16 [B,N,3,H,hd] buffer")
    if kv_len is not None and (kv_len.dtype != torch.int32 or kv_len.numel() != B or not kv_len.is_contiguous()):
        raise B200Error("attention_fwd: kv_len must be a contiguous int32 [B] tensor")
    out = torch.empty(B, N, H * hd, dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty(B, H, N, dtype=torch.float32, device=qkv.device) if need_lse else None
    rc = _lib.load().b200_attn_fwd(qkv.data_ptr(), out.data_ptr(), _ptr(lse), _ptr(kv_len), B, N, H, hd, float(scale),
                                   _stream())
    _lib.check(rc, "b200_attn_fwd")
    _count()
    return out, lse


# --------------------------------------------------------------------------------------------- patch embed helpers
def patchify(img, p, kpad=None):
    """NCHW image (bf16 / fp32) -> [B*gh*gw, Kpad] bf16 patch matrix (k = c*p*p + ky*p + kx)."""
    _need_cuda(img)
    if img.dim() != 4 or not img.is_contiguous():
        raise B200Error("patchify: need a contiguous NCHW image")
    B, C, H, W = img.shape
    K = C * p * p
    kpad = kpad or (K + 7) // 8 * 8
    gh, gw = H // p, W // p
    out = torch.empty(B * gh * gw, kpad, dtype=torch.bfloat16, device=img.device)
    rc = _lib.load().b200_patchify(img.data_ptr(), out.data_ptr(), B, C, H, W, p, kpad, _dt(img), _stream())
    _lib.check(rc, "b200_patchify")
    _count()
    return out


def assemble_tokens(proj, prefix, pos, B, G, n_prefix, D, pos_has_prefix=True):
    """tokens[b] = cat(prefix, proj[b]) + pos (pos_has_prefix) or cat(prefix, proj[b] + pos)  ->  bf16 [B, n_prefix+G, D]."""
    _need_cuda(proj, prefix, pos)
    if proj.dtype != torch.bfloat16 or not proj.is_contiguous() or proj.numel() != B * G * D:
        raise B200Error("assemble_tokens: proj must be contiguous bf16 [B*G, D]")
    npos = (n_prefix + G) if pos_has_prefix else G
    for t, n, name in ((prefix, n_prefix * D, "prefix"), (pos, npos * D, "pos")):
        if t is not None and (not t.is_contiguous() or t.numel() != n):
            raise B200Error(f"assemble_tokens: {name} has {t.numel()} elements, expected {n}")
    tokens = torch.empty(B, n_prefix + G, D, dtype=torch.bfloat16, device=proj.device)
    rc = _lib.load().b200_assemble_tokens(proj.data_ptr(), _ptr(prefix), _ptr(pos), tokens.data_ptr(), B, G, n_prefix, D,
                                          _param_dtype(prefix, pos), int(pos_has_prefix), _stream())
    _lib.check(rc, "b200_assemble_tokens")
    _count()
    return tokens


def colsum(x2d, out=None, accumulate=False):
    """fp32 column sums of a contiguous bf16 [rows, cols] matrix."""
    _need_cuda(x2d, out)
    if x2d.dim() != 2 or x2d.dtype != torch.bfloat16 or not x2d.is_contiguous():
        raise B200Error("colsum: need a contiguous 2-D bf16 tensor")
    rows, cols = x2d.shape
    if out is None:
        out = torch.empty(cols, dtype=torch.float32, device=x2d.device)
        accumulate = False
    rc = _lib.load().b200_colsum(x2d.data_ptr(), out.data_ptr(), rows, cols, int(accumulate), _stream())
    _lib.check(rc, "b200_colsum")
    _count()
    return out
