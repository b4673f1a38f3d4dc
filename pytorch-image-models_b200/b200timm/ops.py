"""Tensor-level wrappers over the C ABI (raw forward / backward kernels, no autograd).

PyTorch is used here only for device memory (caching allocator) and the current CUDA stream; every
numerical operation is a libb200timm kernel.  All functions require CUDA tensors and raise on anything
the kernels do not support (no fallback).
"""
import ctypes

import torch

from . import _lib
from ._lib import (DTYPE_BF16, DTYPE_F32, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_BIAS_ROPE, EPI_DGELU, B200Error,
                   GemmDesc)

__all__ = [
    "layernorm_fwd", "layernorm_bwd", "gemm", "linear_fwd", "linear_dgrad", "linear_wgrad", "convert_f32",
    "attention_fwd", "attention_bwd", "patchify", "assemble_tokens", "colsum", "dwconv_nhwc", "dwconv_nhwc_wgrad", "unpatchify_nhwc", "patchify_nhwc",
    "avgpool_nhwc", "naflex_assemble", "naflex_assemble_bwd", "set_input_norm", "input_norm", "glu_fwd", "glu_bwd", "qk_norm_fwd", "qk_norm_bwd", "grn_fwd", "grn_bwd", "scale_rows_cols", "colsum_mul", "EPI_BIAS", "EPI_BIAS_GELU", "EPI_BIAS_RESIDUAL", "EPI_DGELU", "EPI_BIAS_ROPE", "rope_qk", "RmsEps", "ACTS",
]


def _stream():
    # raw cudaStream_t of torch's current stream on the current device; torch.cuda.current_stream() costs ~13 us of
    # Python per call (device-index / availability checks), this is a direct C call
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


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


# ------------------------------------------------------------------------------------ weight-cache invalidation
_param_epoch = 0       # bumped whenever a kernel wrote parameters through raw pointers (optimizer step, EMA, graph replay)


def wkey(t):
    """Cache key of a derived copy of tensor `t` (bf16 cast, padded / transposed weight ...): storage address, autograd
    version counter AND the process-wide 'parameters written by a kernel' epoch -- raw-pointer writes of the optimizer /
    EMA kernels do not touch `_version` on their own, and a CUDA-graph replay of a training step runs no Python at all."""
    return (t.data_ptr(), t._version, _param_epoch)


def note_params_written(tensors=None):
    """Called after every launch that writes parameters through raw pointers: bumps the tensors' version counters (so
    autograd's saved-tensor checks and the caches keyed on `_version` see the write) and the global epoch."""
    global _param_epoch
    _param_epoch += 1
    if tensors:
        torch._C._increment_version(list(tensors))


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
class RmsEps(float):
    """An eps that marks RMS mode: layernorm_fwd(x, w, None, RmsEps(1e-6)) computes x * rsqrt(mean(x^2) + eps) * w
    (timm RmsNorm) and returns mean = None, which in turn makes layernorm_bwd take the RMS backward.  Carrying the mode
    on the eps keeps every fused call site (BlockFn, LayerNormFn, ...) mode-agnostic."""


def layernorm_fwd(x, weight, bias, eps, out_dtype=None, save_stats=False):
    """y = LayerNorm(x) (or RMSNorm(x) when eps is an RmsEps) over the last dim.  Returns (y, mean, rstd); stats are None
    unless save_stats, mean is always None in RMS mode."""
    _need_cuda(x, weight, bias)
    if isinstance(eps, RmsEps):
        if bias is not None:
            raise B200Error("RMSNorm has no bias")
        return _rmsnorm_fwd(x, weight, eps, out_dtype, save_stats)
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


def _rmsnorm_fwd(x, weight, eps, out_dtype, save_stats):
    cols = x.shape[-1]
    x2 = x.reshape(-1, cols)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    rows = x2.shape[0]
    y = torch.empty(x2.shape, dtype=out_dtype or x.dtype, device=x.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if save_stats else None
    rc = _lib.load().b200_rmsnorm_fwd(x2.data_ptr(), _ptr(weight), y.data_ptr(), _ptr(rstd), rows, cols, float(eps), _dt(x2),
                                      _dt(y), _param_dtype(weight), _stream())
    _lib.check(rc, "b200_rmsnorm_fwd")
    _count()
    return y.view(*x.shape[:-1], cols), None, rstd


def layernorm_bwd(dy, x, weight, mean, rstd, need_param_grads=True, dres=None, dw_out=None, db_out=None):
    """Returns (dx, dweight fp32, dbias fp32).  dres (dtype/shape of x): gradient of the residual branch around
    the norm, added into dx in the same pass.  mean = None selects the RMSNorm backward (dbias is None then)."""
    _need_cuda(dy, x, weight, mean, rstd, dres)
    rms = mean is None
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
        dw = dw_out if dw_out is not None else torch.empty(cols, dtype=torch.float32, device=x.device)
        db = None if rms else (db_out if db_out is not None else torch.empty(cols, dtype=torch.float32, device=x.device))
    if dres is not None:
        dres = dres.reshape(-1, cols)
        if dres.dtype != x2.dtype or dres.shape != x2.shape:
            raise B200Error("layernorm_bwd: dres must match x in dtype and shape")
        if not dres.is_contiguous():
            dres = dres.contiguous()
    if rms:
        rc = lib.b200_rmsnorm_bwd(dy2.data_ptr(), x2.data_ptr(), _ptr(weight), rstd.data_ptr(), _ptr(dres), dx.data_ptr(),
                                  _ptr(dw), ws.data_ptr(), rows, cols, _dt(x2), _dt(dy2), _param_dtype(weight), _stream())
    else:
        rc = lib.b200_layernorm_bwd(dy2.data_ptr(), x2.data_ptr(), _ptr(weight), mean.data_ptr(), rstd.data_ptr(),
                                    _ptr(dres), dx.data_ptr(), _ptr(dw), _ptr(db), ws.data_ptr(), rows, cols, _dt(x2), _dt(dy2),
                                    _param_dtype(weight), _stream())
    _lib.check(rc, "b200_layernorm_bwd")
    _count(2 if need_param_grads else 1)
    return dx.view(x.shape), dw, db


# --------------------------------------------------------------------------------------------- GEMM
def _check_2d_bf16(t, name):
    if t.dim() != 2 or t.dtype != torch.bfloat16 or t.stride(1) != 1:
        raise B200Error(f"{name}: need a 2-D bf16 tensor with unit inner stride, got {tuple(t.shape)} {t.dtype} "
                        f"strides {t.stride()}")


ACTS = {"gelu": 0, "gelu_tanh": 1, "silu": 2}     # b200_gemm_desc.act: the activation of EPI_BIAS_GELU / EPI_DGELU


def gemm(a, b, M, N, K, *, a_mn=False, b_mn=False, epilogue=EPI_BIAS, bias=None, aux=None, gamma=None, out=None,
         out_dtype=torch.bfloat16, accumulate=False, split_k=1, out2=None, row_scale=None, rows_per_scale=0, rope=None,
         act=0, colsum=None, ln_stats_out=None, ln_in=None):
    """out[M,N] = epilogue(sum_k A[m,k] B[n,k]).

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True) bf16; b: [N,K] (b_mn=False) or [K,N] (b_mn=True) bf16.
    ln_stats_out: fp32 [ln_slots(M, N) + 1, M, 2] (ln_stats_buffer) that receives the row statistics of `out` (producer side of
    a folded LayerNorm); ln_in = (that buffer, c1 fp32 [N], eps): consumer side -- `a` is the un-normalised activation, `b` the weight times
    the norm's weight, `bias` the fp32 vector ln_bias . W^T + linear bias (b200_gemm_desc.ln_*).
    """
    _need_cuda(a, b, bias, aux, gamma, out, out2)
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
    if out2 is not None:
        _check_2d_bf16(out2, "out2")
        if tuple(out2.shape) != (M, N):
            raise B200Error("gemm: out2 must be [M,N]")
        d.out2, d.ldo2 = out2.data_ptr(), out2.stride(0)
    if row_scale is not None:
        if row_scale.dtype != torch.float32 or not row_scale.is_contiguous() or not row_scale.is_cuda:
            raise B200Error("gemm: row_scale must be a contiguous fp32 CUDA tensor")
        if rows_per_scale <= 0 or row_scale.numel() * rows_per_scale < M:
            raise B200Error("gemm: row_scale needs rows_per_scale > 0 and ceil(M / rows_per_scale) entries")
        d.row_scale, d.rows_per_scale = row_scale.data_ptr(), int(rows_per_scale)
    if rope is not None:
        emb, tokens, prefix, cols = rope       # ([tokens - prefix, 128] fp32 = [sin | cos], N, prefix tokens, q+k columns)
        _check_rope(emb, tokens, prefix)
        d.rope, d.rope_tokens, d.rope_prefix, d.rope_cols = emb.data_ptr(), int(tokens), int(prefix), int(cols)
    d.M, d.N, d.K = M, N, K
    d.a_mn, d.b_mn = int(a_mn), int(b_mn)
    d.epilogue = epilogue
    d.out_dtype = _dt(out)
    d.bias_dtype = _dt(bias) if bias is not None else DTYPE_BF16
    d.gamma_dtype = _dt(gamma) if gamma is not None else DTYPE_BF16
    d.accumulate = int(accumulate)
    d.split_k = int(split_k)
    d.act = int(act)
    if colsum is not None:     # fp32 [N], zero-initialised by the caller: += column sums of `out` (EPI_DGELU epilogue)
        if colsum.dtype != torch.float32 or colsum.numel() != N or not colsum.is_contiguous() or not colsum.is_cuda:
            raise B200Error("gemm: colsum must be a contiguous fp32 CUDA tensor [N]")
        d.colsum = colsum.data_ptr()
    if ln_stats_out is not None:
        if (ln_stats_out.dtype != torch.float32 or not ln_stats_out.is_contiguous() or not ln_stats_out.is_cuda
                or tuple(ln_stats_out.shape) != (ln_slots(M, N) + 1, M, 2)):
            raise B200Error("gemm: ln_stats_out must be a contiguous fp32 CUDA tensor [ln_slots(M, N) + 1, M, 2]")
        d.ln_stats_out = ln_stats_out.data_ptr()
    if ln_in is not None:
        st, c1, eps = ln_in
        if (st.dtype != torch.float32 or not st.is_contiguous() or st.dim() != 3 or st.shape[0] < 2 or st.shape[1] != M or st.shape[2] != 2
                or c1.dtype != torch.float32 or c1.numel() != N or not c1.is_contiguous()
                or bias is None or bias.dtype != torch.float32):
            raise B200Error("gemm: ln_in needs stats fp32 [slots + 1, M, 2], c1 fp32 [N] and an fp32 bias")
        _need_cuda(st, c1)
        d.ln_stats, d.ln_c1, d.ln_slots, d.ln_eps = st.data_ptr(), c1.data_ptr(), int(st.shape[0]) - 1, float(eps)
        _count()           # the statistics finalisation launch
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


def ln_slots(M, N):
    """Partial-statistic slots a gemm(..., ln_stats_out=...) launch writes for an [M, N] output."""
    return int(_lib.load().b200_gemm_ln_slots(int(M), int(N)))


def ln_stats_buffer(M, N, device):
    """The statistics buffer of a folded LayerNorm over an [M, N] GEMM output: the producer's slots + the finalised slot."""
    return torch.empty(ln_slots(M, N) + 1, M, 2, dtype=torch.float32, device=device)


def linear_fwd(x2d, weight, bias=None, epilogue=EPI_BIAS, aux=None, gamma=None, out=None, out2=None, row_scale=None,
               rows_per_scale=0, rope=None, act=0, ln_stats_out=None, ln_in=None):
    """F.linear on a [M,K] bf16 activation with a fused epilogue (weight: [N,K] bf16)."""
    M, K = x2d.shape
    N = weight.shape[0]
    return gemm(x2d, weight, M, N, K, epilogue=epilogue, bias=bias, aux=aux, gamma=gamma, out=out, out2=out2,
                row_scale=row_scale, rows_per_scale=rows_per_scale, rope=rope, act=act, ln_stats_out=ln_stats_out, ln_in=ln_in)


def _check_rope(emb, tokens, prefix, batch=None):
    """Shared table [tokens - prefix, 128]; with `batch`, also a per-sample one [batch, tokens - prefix, 128]."""
    ok = emb.dtype == torch.float32 and emb.is_cuda and emb.is_contiguous() and emb.shape[-1] == 128 and prefix >= 0
    if ok and emb.dim() == 2:
        ok = emb.shape[0] == tokens - prefix
    elif ok and emb.dim() == 3 and batch is not None:
        ok = emb.shape[0] == batch and emb.shape[1] == tokens - prefix
    else:
        ok = False
    if not ok:
        raise B200Error("rope: the embedding must be a contiguous fp32 CUDA tensor [tokens - prefix, 2 * 64] = [sin | cos] "
                        "(or [B, tokens - prefix, 128] for per-sample tables)")


def rope_qk(qkv, emb, B, N, H, num_prefix, half=False, inverse=False):
    """Rotate the q / k thirds of qkv [B*N, 3*H*64] bf16 IN PLACE (apply_rot_embed_cat, attention.py:249-253);
    inverse=True applies the transposed map (backward of the rotation).  Returns qkv."""
    _need_cuda(qkv, emb)
    if qkv.dtype != torch.bfloat16 or not qkv.is_contiguous() or qkv.numel() != B * N * 3 * H * 64:
        raise B200Error("rope_qk: qkv must be a contiguous bf16 [B*N, 3*H*64] buffer (head_dim 64)")
    _check_rope(emb, N, num_prefix, batch=B)
    rc = _lib.load().b200_rope_qk(qkv.data_ptr(), emb.data_ptr(), B, N, H, int(num_prefix), int(half), int(inverse),
                                  int(emb.dim() == 3), _stream())
    _lib.check(rc, "b200_rope_qk")
    _count()
    return qkv


def dgrad_colsum_fused(M, K):
    """True when linear_dgrad(..., epilogue=EPI_DGELU, colsum=...) can fold the column sums of its output into the epilogue."""
    return M >= 256 and K > 128


def linear_dgrad(dy2d, weight, epilogue=EPI_BIAS, aux=None, act=0, colsum=None):
    """dX[M,K] = dY[M,N] @ W[N,K]  (W consumed MN-major in place; optional act' epilogue with aux = pre-activation).
    colsum (EPI_DGELU, dgrad_colsum_fused shapes): zeroed fp32 [K] that receives the column sums of dX -- the bias gradient
    of the Linear that produced the pre-activation."""
    M, N = dy2d.shape
    K = weight.shape[1]
    return gemm(dy2d, weight, M, K, N, b_mn=True, epilogue=epilogue, aux=aux, act=act, colsum=colsum)


def _wgrad_split(rows_out, cols_out, k):
    """Split-K factor that fills the SMs: tiles of 128 x 256 (or 128 x 128 when cols <= 128)."""
    tiles = ((rows_out + 127) // 128) * ((cols_out + 255) // 256 if cols_out > 128 else 1)
    sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    split = max(1, min(sms // max(tiles, 1), (k + 4095) // 4096))
    return split


_WGRAD_NSPLIT = __import__("os").environ.get("B200_WGRAD_NSPLIT", "1")[:1] != "0"


def _wgrad_parts(rows_out, cols_out, k):
    """Output-row chunks for one weight gradient.  The pair kernel covers [rows_out, cols_out] with 256 x 256 tiles, each
    split along K so that the tile-pairs fill the 74 CTA pairs of the GPU; a shape whose tile count divides 74 badly (ViT-L's
    fused qkv projection [3072, 1024]: 48 tiles, no split -> 48 of 74 pairs busy) runs as several launches over row chunks
    that do (2 x [1536, 1024]: 24 tiles, split 3 -> 72 pairs each; measured 191 -> 166 us at 36928 rows).  Chunks that would
    need a deep split are not taken (ViT-B's [2304, 768] as 3 x [768, 768] with split 8 measured 147 us against 139 us for the
    single launch at 54 pairs: the reduce-add traffic of 8 partial tiles outweighs the idle SMs).  Returns [(row0, rows,
    split_k)] or None."""
    sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    pairs = max(1, sms // 2)

    def plan(rows):
        tiles = ((rows + 255) // 256) * ((cols_out + 255) // 256)
        split = max(1, min(pairs // max(tiles, 1), (k + 4095) // 4096))
        return tiles, split
    tiles, split = plan(rows_out)
    if not _WGRAD_NSPLIT or rows_out < 512 or cols_out <= 128 or k < 256 or tiles * split >= 0.85 * pairs:
        return None
    best = None
    for parts in (2, 3, 4):
        if rows_out % (parts * 256):
            continue
        t, sp = plan(rows_out // parts)
        util = t * sp / pairs
        if util <= 1.0 and sp <= 4 and (best is None or util > best[0]):
            best = (util, parts, sp)
    if best is None or best[0] < tiles * split / pairs + 0.15:
        return None
    _, parts, sp = best
    rows = rows_out // parts
    return [(i * rows, rows, max(2, sp)) for i in range(parts)]


def linear_wgrad(dy2d, x2d, out=None):
    """dW[N,K] = dY[M,N]^T @ X[M,K] in fp32 (both operands MN-major in place, split-K through TMA reduce-add).
    `out`: optional ZERO-INITIALISED (or to-be-accumulated-into) fp32 [N,K] buffer."""
    M, N = dy2d.shape
    K = x2d.shape[1]
    ws = out if out is not None else torch.zeros(N, K, dtype=torch.float32, device=dy2d.device)
    parts = _wgrad_parts(N, K, M) if M >= 256 and ws.is_contiguous() else None
    if parts is None:
        gemm(dy2d, x2d, N, K, M, a_mn=True, b_mn=True, out=ws, split_k=max(2, _wgrad_split(N, K, M)))
    else:
        for r0, rows, split in parts:      # column ranges of dY (views: same row pitch) -> row ranges of dW
            gemm(dy2d[:, r0:r0 + rows], x2d, rows, K, M, a_mn=True, b_mn=True, out=ws[r0:r0 + rows], split_k=split)
    return ws


def convert_f32(src, dst, accumulate=False):
    """dst (bf16 / fp32) (+)= src (fp32), element count a multiple of 4."""
    _need_cuda(src, dst)
    if src.dtype != torch.float32 or not src.is_contiguous() or not dst.is_contiguous() or src.numel() != dst.numel():
        raise B200Error("convert_f32: need contiguous tensors of equal size, fp32 source")
    rc = _lib.load().b200_convert_f32(src.data_ptr(), dst.data_ptr(), src.numel(), _dt(dst), int(accumulate), _stream())
    _lib.check(rc, "b200_convert_f32")
    _count()
    return dst


# --------------------------------------------------------------------------------------------- attention
def _attn_bias_args(mask, B, N, H, who):
    """(ptr, batch stride, head stride, row pitch) of an additive attention bias: contiguous fp32 [B or 1, H or 1, N, ld]."""
    if (mask.dim() != 4 or not mask.is_contiguous() or mask.shape[0] not in (1, B) or mask.shape[1] not in (1, H)
            or mask.shape[2] != N or mask.shape[3] < N or mask.shape[3] % 4):
        raise B200Error(f"{who}: the attention bias must be contiguous fp32 [B or 1, H or 1, N, ld >= N (multiple of 4)], "
                        f"got {tuple(mask.shape)} for B={B} H={H} N={N}")
    ld = mask.shape[3]
    return mask.data_ptr(), (mask.shape[1] * N * ld if mask.shape[0] > 1 else 0), (N * ld if mask.shape[1] > 1 else 0), ld


def attention_fwd(qkv, B, N, H, hd, scale, kv_len=None, need_lse=False):
    """qkv: bf16 [B*N, 3*H*hd] (or [B,N,3*H*hd]) contiguous.  Returns (out [B,N,H*hd] bf16, lse [B,H,N] | None).
    kv_len: None, int32 [B] valid-key counts (prefix key-padding mask), or an fp32 4-D additive bias (causal / general masks;
    layers.attn_mask_operand builds either from the reference's (attn_mask, is_causal))."""
    _need_cuda(qkv, kv_len)
    if qkv.dtype != torch.bfloat16 or not qkv.is_contiguous() or qkv.numel() != B * N * 3 * H * hd:
        raise B200Error("attention_fwd: qkv must be a contiguous bf16 [B,N,3,H,hd] buffer")
    out = torch.empty(B, N, H * hd, dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty(B, H, N, dtype=torch.float32, device=qkv.device) if need_lse else None
    if kv_len is not None and kv_len.dtype == torch.float32:
        ptr, bs, hs, ld = _attn_bias_args(kv_len, B, N, H, "attention_fwd")
        rc = _lib.load().b200_attn_fwd_bias(qkv.data_ptr(), out.data_ptr(), _ptr(lse), ptr, bs, hs, ld, B, N, H, hd, float(scale),
                                            _stream())
        _lib.check(rc, "b200_attn_fwd_bias")
        _count()
        return out, lse
    if kv_len is not None and (kv_len.dtype != torch.int32 or kv_len.numel() != B or not kv_len.is_contiguous()):
        raise B200Error("attention_fwd: kv_len must be a contiguous int32 [B] tensor")
    rc = _lib.load().b200_attn_fwd(qkv.data_ptr(), out.data_ptr(), _ptr(lse), _ptr(kv_len), B, N, H, hd, float(scale),
                                   _stream())
    _lib.check(rc, "b200_attn_fwd")
    _count()
    return out, lse


def attn_bwd_bias_fused(N, hd=64):
    """True when attention_bwd(..., dbias=...) can reduce the qkv bias gradient from its output tiles."""
    import os
    return N <= 208 and hd == 64 and os.environ.get("B200_ATTN_BWD_V1", "0")[:1] != "1"


def attention_bwd(qkv, out, dout, lse, B, N, H, hd, scale, kv_len=None, dbias=None):
    """dqkv [B,N,3*H*hd] bf16 from dout [B,N,H*hd]; recomputes P from qkv and the saved log-sum-exp.  dbias (attn_bwd_bias_fused
    shapes): zeroed fp32 [3*H*hd] that receives the column sums of dqkv (the qkv projection's bias gradient)."""
    _need_cuda(qkv, out, dout, lse, kv_len)
    for t, n, name in ((qkv, B * N * 3 * H * hd, "qkv"), (out, B * N * H * hd, "out"), (dout, B * N * H * hd, "dout")):
        if t.dtype != torch.bfloat16 or not t.is_contiguous() or t.numel() != n:
            raise B200Error(f"attention_bwd: {name} must be a contiguous bf16 buffer of {n} elements")
    if lse.dtype != torch.float32 or not lse.is_contiguous() or lse.numel() != B * H * N:
        raise B200Error("attention_bwd: lse must be contiguous fp32 [B,H,N]")
    dqkv = torch.empty_like(qkv)
    lib = _lib.load()
    delta = torch.empty(lib.b200_attn_bwd_ws_floats_hd(B, N, H, hd), dtype=torch.float32, device=qkv.device)   # delta (+ fp32 dQ)
    if dbias is not None and (dbias.dtype != torch.float32 or dbias.numel() != 3 * H * hd or not dbias.is_contiguous()):
        raise B200Error("attention_bwd: dbias must be a contiguous fp32 [3*H*hd] tensor")
    if kv_len is not None and kv_len.dtype == torch.float32:
        ptr, bs, hs, ld = _attn_bias_args(kv_len, B, N, H, "attention_bwd")
        rc = lib.b200_attn_bwd_bias(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), delta.data_ptr(),
                                    dqkv.data_ptr(), ptr, bs, hs, ld, B, N, H, hd, float(scale), _ptr(dbias), _stream())
        _lib.check(rc, "b200_attn_bwd_bias")
        _count(2)
        return dqkv
    rc = lib.b200_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), delta.data_ptr(),
                                   dqkv.data_ptr(), _ptr(kv_len), B, N, H, hd, float(scale), _ptr(dbias), _stream())
    _lib.check(rc, "b200_attn_bwd")
    _count(2)
    return dqkv


# --------------------------------------------------------------------------------------------- patch embed helpers
_input_norm = None      # (mean * 255, std * 255) as ctypes float arrays: the loader's normalisation for uint8 images


def set_input_norm(mean=None, std=None):
    """Normalisation folded into the patch gather for uint8 images: PrefetchLoader's `(x - mean * 255) / (std * 255)`
    (timm/data/loader.py:106-109,139) with `mean` / `std` per channel in [0, 1] units.  None switches it off."""
    global _input_norm
    if mean is None:
        _input_norm = None
        return
    mean, std = [float(m) * 255 for m in mean], [float(s) * 255 for s in std]
    if len(mean) != len(std) or not 1 <= len(mean) <= 8 or any(s == 0 for s in std):
        raise B200Error("set_input_norm: mean / std need 1..8 channels and non-zero std")
    if _input_norm is not None and (list(_input_norm[0]) != [ctypes.c_float(m).value for m in mean]
                                    or list(_input_norm[1]) != [ctypes.c_float(s).value for s in std]):
        # the setting is process-wide (the patch gather of EVERY uint8 batch uses it): a second loader with other
        # statistics would silently re-normalise the first one's batches
        raise B200Error("set_input_norm: a different uint8 normalisation is already active in this process; call "
                        "set_input_norm(None) first (one mean / std per process)")
    # the loader builds these as fp32 tensors: round to fp32 the same way
    _input_norm = ((ctypes.c_float * len(mean))(*mean), (ctypes.c_float * len(std))(*std))


def input_norm():
    return _input_norm


def patchify(img, p, kpad=None):
    """NCHW image (bf16 / fp32, or uint8 with set_input_norm) -> [B*gh*gw, Kpad] bf16 patch matrix (k = c*p*p + ky*p + kx)."""
    _need_cuda(img)
    if img.dim() != 4 or not img.is_contiguous():
        raise B200Error("patchify: need a contiguous NCHW image")
    B, C, H, W = img.shape
    K = C * p * p
    kpad = kpad or (K + 7) // 8 * 8
    gh, gw = H // p, W // p
    out = torch.empty(B * gh * gw, kpad, dtype=torch.bfloat16, device=img.device)
    if img.dtype == torch.uint8:
        if _input_norm is None or len(_input_norm[0]) != C:
            raise B200Error("uint8 images need ops.set_input_norm(mean, std) with one entry per channel")
        rc = _lib.load().b200_patchify_u8(img.data_ptr(), out.data_ptr(), B, C, H, W, p, kpad, _input_norm[0],
                                          _input_norm[1], _stream())
        _lib.check(rc, "b200_patchify_u8")
        _count()
        return out
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


# --------------------------------------------------------------------------------------------- ConvNeXt (NHWC)
def dwconv_weight_t(weight):
    """nn.Conv2d depthwise weight [C,1,k,k] -> the [k,k,C] layout b200_dwconv_nhwc consumes (do once, cache)."""
    return weight.detach().reshape(weight.shape[0], -1).t().contiguous().view(weight.shape[-2], weight.shape[-1], -1)


def dwconv_nhwc(x, weight_t, bias, residual=None):
    """Depthwise k x k conv (stride 1, pad k//2) on a contiguous bf16 [B,H,W,C] tensor; weight_t [k,k,C];
    optional residual (same shape as x) added to the output."""
    _need_cuda(x, weight_t, bias, residual)
    if residual is not None and (residual.dtype != torch.bfloat16 or residual.shape != x.shape or not residual.is_contiguous()):
        raise B200Error("dwconv_nhwc: residual must be a contiguous bf16 tensor shaped like x")
    if x.dim() != 4 or x.dtype != torch.bfloat16 or not x.is_contiguous():
        raise B200Error("dwconv_nhwc: need a contiguous bf16 [B,H,W,C] tensor")
    B, H, W, C = x.shape
    k = weight_t.shape[0]
    if tuple(weight_t.shape) != (k, k, C) or weight_t.dtype != torch.bfloat16 or not weight_t.is_contiguous():
        raise B200Error("dwconv_nhwc: weight_t must be contiguous bf16 [k,k,C] (see dwconv_weight_t)")
    y = torch.empty_like(x)
    rc = _lib.load().b200_dwconv_nhwc(x.data_ptr(), weight_t.data_ptr(), _ptr(bias), _ptr(residual), y.data_ptr(), B, H, W, C, k,
                                      _dt(bias) if bias is not None else DTYPE_BF16, _stream())
    _lib.check(rc, "b200_dwconv_nhwc")
    _count()
    return y


def dwconv_ln_fusable(C, k):
    """True when dwconv_nhwc(..., ln_stats=True) can produce the statistics of a folded LayerNorm (the TMA-staged kernel)."""
    return C % 64 == 0 and k == 7


def dwconv_nhwc_ln(x, weight_t, bias):
    """dwconv_nhwc that also returns the partial LayerNorm statistics of its output rows: (y [B,H,W,C], stats fp32
    [C / 64 + 1, B*H*W, 2]) -- the `ln_in` buffer of the GEMM that follows the block's norm (b200_dwconv_nhwc_ln)."""
    _need_cuda(x, weight_t, bias)
    if x.dim() != 4 or x.dtype != torch.bfloat16 or not x.is_contiguous():
        raise B200Error("dwconv_nhwc_ln: need a contiguous bf16 [B,H,W,C] tensor")
    B, H, W, C = x.shape
    k = weight_t.shape[0]
    if tuple(weight_t.shape) != (k, k, C) or weight_t.dtype != torch.bfloat16 or not weight_t.is_contiguous():
        raise B200Error("dwconv_nhwc_ln: weight_t must be contiguous bf16 [k,k,C] (see dwconv_weight_t)")
    if not dwconv_ln_fusable(C, k):
        raise B200Error("dwconv_nhwc_ln: needs C % 64 == 0 and a 7 x 7 kernel")
    y = torch.empty_like(x)
    stats = torch.empty(C // 64 + 1, B * H * W, 2, dtype=torch.float32, device=x.device)
    rc = _lib.load().b200_dwconv_nhwc_ln(x.data_ptr(), weight_t.data_ptr(), _ptr(bias), y.data_ptr(), stats.data_ptr(), B, H, W, C, k,
                                         _dt(bias) if bias is not None else DTYPE_BF16, _stream())
    _lib.check(rc, "b200_dwconv_nhwc_ln")
    _count()
    return y, stats


def dwconv_nhwc_wgrad(x, dy, k, out=None):
    """fp32 [k,k,C] gradient of the depthwise weight (transposed layout); `out` must be zero-initialised."""
    _need_cuda(x, dy, out)
    for t in (x, dy):
        if t.dim() != 4 or t.dtype != torch.bfloat16 or not t.is_contiguous() or t.shape != x.shape:
            raise B200Error("dwconv_nhwc_wgrad: need contiguous bf16 [B,H,W,C] tensors of equal shape")
    B, H, W, C = x.shape
    if out is None:
        out = torch.zeros(k, k, C, dtype=torch.float32, device=x.device)
    rc = _lib.load().b200_dwconv_nhwc_wgrad(x.data_ptr(), dy.data_ptr(), out.data_ptr(), B, H, W, C, k, _stream())
    _lib.check(rc, "b200_dwconv_nhwc_wgrad")
    _count()
    return out


def unpatchify_nhwc(patches, B, H, W, C, p):
    """Inverse of patchify_nhwc: [B*(H/p)*(W/p), p*p*C] -> [B,H,W,C]."""
    _need_cuda(patches)
    if patches.dtype != torch.bfloat16 or not patches.is_contiguous() or patches.numel() != B * H * W * C:
        raise B200Error("unpatchify_nhwc: need a contiguous bf16 patch matrix of B*H*W*C elements")
    out = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=patches.device)
    rc = _lib.load().b200_unpatchify_nhwc(patches.data_ptr(), out.data_ptr(), B, H, W, C, p, _stream())
    _lib.check(rc, "b200_unpatchify_nhwc")
    _count()
    return out


def patchify_nhwc(x, p):
    """[B,H,W,C] bf16 -> [B*(H/p)*(W/p), p*p*C] (k = (ky*p + kx)*C + c)."""
    _need_cuda(x)
    if x.dim() != 4 or x.dtype != torch.bfloat16 or not x.is_contiguous():
        raise B200Error("patchify_nhwc: need a contiguous bf16 [B,H,W,C] tensor")
    B, H, W, C = x.shape
    out = torch.empty(B * (H // p) * (W // p), p * p * C, dtype=torch.bfloat16, device=x.device)
    rc = _lib.load().b200_patchify_nhwc(x.data_ptr(), out.data_ptr(), B, H, W, C, p, _stream())
    _lib.check(rc, "b200_patchify_nhwc")
    _count()
    return out


def avgpool_nhwc(x):
    """[B,HW,C] (or [B,H,W,C]) bf16 -> [B,C] mean over the spatial positions."""
    _need_cuda(x)
    if x.dtype != torch.bfloat16 or not x.is_contiguous() or x.dim() not in (3, 4):
        raise B200Error("avgpool_nhwc: need a contiguous bf16 [B,HW,C] tensor")
    B, C = x.shape[0], x.shape[-1]
    out = torch.empty(B, C, dtype=torch.bfloat16, device=x.device)
    rc = _lib.load().b200_avgpool_nhwc(x.data_ptr(), out.data_ptr(), B, x.numel() // (B * C), C, _stream())
    _lib.check(rc, "b200_avgpool_nhwc")
    _count()
    return out


# --------------------------------------------------------------------------------------------- q / k normalisation
def qk_norm_fwd(qkv, T, H, qw, qb, kw, kb, eps, save=False):
    """Normalises the q and k rows of qkv (bf16 [T, 3*H*64]) IN PLACE (Attention.q_norm / k_norm over head_dim 64,
    timm/layers/attention.py:117); eps may be an RmsEps.  save=True returns (raw q/k rows bf16 [T, 2, H, 64], mean, rstd)."""
    _need_cuda(qkv, qw, qb, kw, kb)
    rms = isinstance(eps, RmsEps)
    if qkv.dtype != torch.bfloat16 or not qkv.is_contiguous() or qkv.numel() != T * 3 * H * 64:
        raise B200Error("qk_norm_fwd: qkv must be a contiguous bf16 [T, 3, H, 64] buffer (head_dim 64)")
    for w in (qw, kw):
        if w.numel() != 64 or not w.is_contiguous():
            raise B200Error("qk_norm_fwd: norm weights must be contiguous [64]")
    dt = _param_dtype(qw, qb, kw, kb)
    raw = mean = rstd = None
    if save:
        raw = torch.empty(T, 2, H, 64, dtype=torch.bfloat16, device=qkv.device)
        rstd = torch.empty(T * 2 * H, dtype=torch.float32, device=qkv.device)
        mean = None if rms else torch.empty(T * 2 * H, dtype=torch.float32, device=qkv.device)
    rc = _lib.load().b200_qk_norm_fwd(qkv.data_ptr(), _ptr(raw), _ptr(mean), _ptr(rstd), qw.data_ptr(), _ptr(qb),
                                      kw.data_ptr(), _ptr(kb), dt, T, H, 64, float(eps), int(rms), _stream())
    _lib.check(rc, "b200_qk_norm_fwd")
    _count()
    return raw, mean, rstd


def qk_norm_bwd(dqkv, raw, mean, rstd, qw, kw, T, H):
    """In place on dqkv: gradients w.r.t. the normalised q / k rows -> w.r.t. the raw rows.  Returns fp32 [4, 64]:
    dq_weight, dq_bias, dk_weight, dk_bias.  mean=None selects RMSNorm."""
    _need_cuda(dqkv, raw, mean, rstd, qw, kw)
    if dqkv.dtype != torch.bfloat16 or not dqkv.is_contiguous() or dqkv.numel() != T * 3 * H * 64:
        raise B200Error("qk_norm_bwd: dqkv must be a contiguous bf16 [T, 3, H, 64] buffer")
    dparams = torch.zeros(4, 64, dtype=torch.float32, device=dqkv.device)
    rc = _lib.load().b200_qk_norm_bwd(dqkv.data_ptr(), raw.data_ptr(), _ptr(mean), rstd.data_ptr(), qw.data_ptr(),
                                      kw.data_ptr(), _param_dtype(qw, kw), dparams.data_ptr(), T, H, 64, int(mean is None),
                                      _stream())
    _lib.check(rc, "b200_qk_norm_bwd")
    _count()
    return dparams


# --------------------------------------------------------------------------------------------- Global Response Norm
def grn_fwd(h, weight, bias, B, HW, eps=1e-6):
    """GlobalResponseNorm (timm/layers/grn.py:46-49) IN PLACE on channels-last rows h bf16 [B*HW, C].  Returns the saved
    (gsq fp32 [B, C], scale fp32 [B, C]) pair the backward needs."""
    _need_cuda(h, weight, bias)
    C = h.shape[-1]
    if h.dtype != torch.bfloat16 or not h.is_contiguous() or h.numel() != B * HW * C or C % 8:
        raise B200Error("grn_fwd: h must be contiguous bf16 [B*HW, C] with C % 8 == 0")
    gsq = torch.zeros(B, C, dtype=torch.float32, device=h.device)
    scale = torch.empty(B, C, dtype=torch.float32, device=h.device)
    rc = _lib.load().b200_grn_fwd(h.data_ptr(), gsq.data_ptr(), scale.data_ptr(), weight.data_ptr(), _ptr(bias),
                                  _param_dtype(weight, bias), B, HW, C, float(eps), _stream())
    _lib.check(rc, "b200_grn_fwd")
    _count(3)
    return gsq, scale


def grn_bwd(dh2, pre, gsq, scale, weight, B, HW, eps=1e-6, dweight=None):
    """Returns (dpre bf16 [B*HW, C] = gradient w.r.t. the fc1 pre-activation (GRN + GELU backward), dweight fp32 [C])."""
    _need_cuda(dh2, pre, gsq, scale, weight)
    C = pre.shape[-1]
    for t, name in ((dh2, "dh2"), (pre, "pre")):
        if t.dtype != torch.bfloat16 or not t.is_contiguous() or t.numel() != B * HW * C:
            raise B200Error(f"grn_bwd: {name} must be contiguous bf16 [B*HW, C]")
    ws = torch.zeros(2, B, C, dtype=torch.float32, device=pre.device)
    if dweight is None:
        dweight = torch.zeros(C, dtype=torch.float32, device=pre.device)
    dpre = torch.empty_like(pre)
    rc = _lib.load().b200_grn_bwd(dh2.data_ptr(), pre.data_ptr(), gsq.data_ptr(), scale.data_ptr(), weight.data_ptr(),
                                  _param_dtype(weight), ws[0].data_ptr(), ws[1].data_ptr(), dweight.data_ptr(),
                                  dpre.data_ptr(), B, HW, C, float(eps), _stream())
    _lib.check(rc, "b200_grn_bwd")
    _count(3)
    return dpre, dweight


# --------------------------------------------------------------------------------------------- gated MLPs
GLU_ACTS = {"silu": 0, "sigmoid": 1}


def glu_fwd(h, gate_first, act="silu"):
    """h bf16 [M, 2H] -> act(gate) * value, bf16 [M, H] (GluMlp / SwiGLU, timm/layers/mlp.py:106-108,156-158)."""
    _need_cuda(h)
    if h.dtype != torch.bfloat16 or h.dim() != 2 or not h.is_contiguous() or h.shape[1] % 16:
        raise B200Error("glu_fwd: h must be contiguous bf16 [M, 2H] with H % 8 == 0")
    M, H = h.shape[0], h.shape[1] // 2
    out = torch.empty(M, H, dtype=torch.bfloat16, device=h.device)
    rc = _lib.load().b200_glu_fwd(h.data_ptr(), out.data_ptr(), M, H, int(gate_first), GLU_ACTS[act], _stream())
    _lib.check(rc, "b200_glu_fwd")
    _count()
    return out


def glu_bwd(h, dout, gate_first, act="silu"):
    """Gradient of glu_fwd with respect to h: bf16 [M, 2H]."""
    _need_cuda(h, dout)
    M, H = h.shape[0], h.shape[1] // 2
    if dout.dtype != torch.bfloat16 or not dout.is_contiguous() or tuple(dout.shape) != (M, H):
        raise B200Error("glu_bwd: dout must be contiguous bf16 [M, H]")
    dh = torch.empty_like(h)
    rc = _lib.load().b200_glu_bwd(h.data_ptr(), dout.data_ptr(), dh.data_ptr(), M, H, int(gate_first), GLU_ACTS[act], _stream())
    _lib.check(rc, "b200_glu_bwd")
    _count()
    return dh


def patch_embed_supported(img, p, D):
    """True when patch_embed_fwd can read `img` (bf16 NCHW, contiguous) directly: the im2col-free path."""
    if img.dtype != torch.bfloat16 or img.dim() != 4 or not img.is_contiguous():
        return False
    H, W = img.shape[-2:]
    return (p in (8, 16, 32, 64) and H % p == 0 and W % p == 0 and W // p <= 128 and W % 8 == 0 and D % 8 == 0
            and (img.shape[1] * p * p) % 64 == 0 and img.data_ptr() % 16 == 0)


def patch_embed_fwd(img, weight2d, bias, pos2d, pos_off, out_tokens, row_off, p):
    """out[b, row_off + g] = patch(b, g) . weight2d^T + bias (+ pos2d[pos_off + g]) -> bf16 [B, out_tokens, D]; rows below
    row_off are NOT written.  The image is the GEMM's A operand (5-D tensor map): no patch matrix is materialised."""
    _need_cuda(img, weight2d, bias, pos2d)
    B, C, H, W = img.shape
    D, K = weight2d.shape
    if weight2d.dtype != torch.bfloat16 or not weight2d.is_contiguous() or K != C * p * p:
        raise B200Error("patch_embed_fwd: weight must be contiguous bf16 [D, C*p*p]")
    if pos2d is not None and (pos2d.dtype != torch.bfloat16 or not pos2d.is_contiguous() or pos2d.dim() != 2 or pos2d.shape[1] != D):
        raise B200Error("patch_embed_fwd: pos must be contiguous bf16 [rows, D]")
    out = torch.empty(B, out_tokens, D, dtype=torch.bfloat16, device=img.device)
    rc = _lib.load().b200_patch_embed_fwd(img.data_ptr(), weight2d.data_ptr(), _ptr(bias), _dt(bias) if bias is not None else DTYPE_BF16,
                                          _ptr(pos2d), 0 if pos2d is None else pos2d.shape[0], int(pos_off), out.data_ptr(), B, C,
                                          H, W, int(p), D, int(out_tokens), int(row_off), _stream())
    _lib.check(rc, "b200_patch_embed_fwd")
    _count()
    return out


def batch_patchify(img, p, channels_last=True, kpad=None):
    """batch_patchify (timm/models/naflexvit.py:159-201, pad=True) as the gather feeding the patch projection: NCHW image
    (bf16 / fp32, or uint8 with set_input_norm) -> (bf16 [B * gh * gw, Kpad], (gh, gw)); flat patch layout P-P-C
    (channels_last) or C-P-P; ragged edges zero-padded."""
    _need_cuda(img)
    if img.dim() != 4 or not img.is_contiguous():
        raise B200Error("batch_patchify: need a contiguous NCHW image")
    B, C, H, W = img.shape
    K = C * p * p
    kpad = kpad or (K + 7) // 8 * 8
    gh, gw = -(-H // p), -(-W // p)
    out = torch.empty(B * gh * gw, kpad, dtype=torch.bfloat16, device=img.device)
    mean = std = None
    if img.dtype == torch.uint8:
        if _input_norm is None or len(_input_norm[0]) != C:
            raise B200Error("uint8 images need ops.set_input_norm(mean, std) with one entry per channel")
        mean, std, dt = _input_norm[0], _input_norm[1], _lib.DTYPE_U8
    else:
        dt = _dt(img)
    rc = _lib.load().b200_batch_patchify(img.data_ptr(), dt, out.data_ptr(), B, C, H, W, p, kpad, int(channels_last), mean, std,
                                         _stream())
    _lib.check(rc, "b200_batch_patchify")
    _count()
    return out, (gh, gw)


def patches_u8(patches, channels, channels_last=True, kpad=None):
    """NaFlexPrefetchLoader's normalisation (timm/data/naflex_loader.py:126-161) of the dataset's uint8 patches
    [..., K] -> bf16 [rows, Kpad] ((v - mean * 255) / (std * 255), set_input_norm), fused with the cast and the K padding."""
    _need_cuda(patches)
    if patches.dtype != torch.uint8 or not patches.is_contiguous():
        raise B200Error("patches_u8: need a contiguous uint8 tensor")
    if _input_norm is None or len(_input_norm[0]) != channels:
        raise B200Error("uint8 patches need ops.set_input_norm(mean, std) with one entry per channel")
    K = patches.shape[-1] if patches.dim() == 3 else patches[0, 0].numel()
    rows = patches.numel() // K
    kpad = kpad or (K + 7) // 8 * 8
    out = torch.empty(rows, kpad, dtype=torch.bfloat16, device=patches.device)
    rc = _lib.load().b200_patches_u8(patches.data_ptr(), out.data_ptr(), rows, K, kpad, channels, int(channels_last),
                                     _input_norm[0], _input_norm[1], _stream())
    _lib.check(rc, "b200_patches_u8")
    _count()
    return out


def act_bwd(dy, pre, act=0):
    """dy * act'(pre) elementwise (bf16, contiguous, equal shapes): Mlp activation backward when an inner norm separates
    the activation from fc2 (otherwise it rides in the fc2-dgrad GEMM epilogue)."""
    _need_cuda(dy, pre)
    if dy.dtype != torch.bfloat16 or pre.dtype != torch.bfloat16 or dy.shape != pre.shape or not dy.is_contiguous() \
            or not pre.is_contiguous() or dy.numel() % 8:
        raise B200Error("act_bwd: need two contiguous bf16 tensors of equal shape, element count a multiple of 8")
    out = torch.empty_like(pre)
    rc = _lib.load().b200_act_bwd(dy.data_ptr(), pre.data_ptr(), out.data_ptr(), dy.numel(), int(act), _stream())
    _lib.check(rc, "b200_act_bwd")
    _count()
    return out


# --------------------------------------------------------------------------------------------- NaFlex
def naflex_assemble(proj, prefix, table, pos_off, pos_len, B, N, n_prefix, D):
    """tokens [B, n_prefix+N, D] = cat(prefix, proj + per-sample interpolated position rows)."""
    _need_cuda(proj, prefix, table, pos_off, pos_len)
    if proj.dtype != torch.bfloat16 or not proj.is_contiguous() or proj.numel() != B * N * D:
        raise B200Error("naflex_assemble: proj must be contiguous bf16 [B*N, D]")
    if table is not None and (table.dtype != torch.bfloat16 or not table.is_contiguous() or table.shape[-1] != D):
        raise B200Error("naflex_assemble: table must be contiguous bf16 [T, D]")
    for t in (pos_off, pos_len):
        if t is not None and (t.dtype != torch.int32 or t.numel() != B or not t.is_contiguous()):
            raise B200Error("naflex_assemble: pos_off / pos_len must be contiguous int32 [B]")
    if prefix is not None and (not prefix.is_contiguous() or prefix.numel() != n_prefix * D):
        raise B200Error("naflex_assemble: prefix must be contiguous [n_prefix, D]")
    tokens = torch.empty(B, n_prefix + N, D, dtype=torch.bfloat16, device=proj.device)
    rc = _lib.load().b200_naflex_assemble(proj.data_ptr(), _ptr(prefix), _ptr(table), _ptr(pos_off), _ptr(pos_len),
                                          tokens.data_ptr(), B, N, n_prefix, D,
                                          _dt(prefix) if prefix is not None else DTYPE_BF16, _stream())
    _lib.check(rc, "b200_naflex_assemble")
    _count()
    return tokens


def naflex_assemble_bwd(dtokens, pos_off, pos_len, table_rows, B, N, n_prefix, D):
    """Backward of naflex_assemble: (dproj bf16 [B*N, D], dprefix fp32 [n_prefix, D] or None, dtable fp32 [T, D] or None)."""
    _need_cuda(dtokens, pos_off, pos_len)
    if dtokens.dtype != torch.bfloat16 or not dtokens.is_contiguous() or dtokens.numel() != B * (N + n_prefix) * D:
        raise B200Error("naflex_assemble_bwd: dtokens must be contiguous bf16 [B, n_prefix+N, D]")
    dproj = torch.empty(B * N, D, dtype=torch.bfloat16, device=dtokens.device)
    dprefix = torch.zeros(n_prefix, D, dtype=torch.float32, device=dtokens.device) if n_prefix else None
    dtable = torch.zeros(table_rows, D, dtype=torch.float32, device=dtokens.device) if table_rows else None
    rc = _lib.load().b200_naflex_assemble_bwd(dtokens.data_ptr(), dproj.data_ptr(), _ptr(dprefix), _ptr(dtable),
                                              _ptr(pos_off), _ptr(pos_len), B, N, n_prefix, D, _stream())
    _lib.check(rc, "b200_naflex_assemble_bwd")
    _count()
    return dproj, dprefix, dtable


# --------------------------------------------------------------------------------------------- LayerScale / DropPath backward
def scale_rows_cols(x2d, col_scale=None, row_scale=None, rows_per_scale=0):
    """x * col_scale[c] * row_scale[r // rows_per_scale] -> bf16 (either scale may be None)."""
    _need_cuda(x2d, col_scale, row_scale)
    if x2d.dim() != 2 or x2d.dtype != torch.bfloat16 or not x2d.is_contiguous():
        raise B200Error("scale_rows_cols: need a contiguous 2-D bf16 tensor")
    rows, cols = x2d.shape
    out = torch.empty_like(x2d)
    rc = _lib.load().b200_scale_rows_cols(x2d.data_ptr(), _ptr(col_scale), _ptr(row_scale), out.data_ptr(), rows, cols,
                                          int(rows_per_scale), _dt(col_scale) if col_scale is not None else DTYPE_BF16,
                                          _stream())
    _lib.check(rc, "b200_scale_rows_cols")
    _count()
    return out


def colsum_mul(a2d, b2d, row_scale=None, rows_per_scale=0, out=None, accumulate=False):
    """fp32 [cols]: sum_r a[r,c] * b[r,c] * row_scale[r // rows_per_scale]."""
    _need_cuda(a2d, b2d, row_scale, out)
    for t in (a2d, b2d):
        if t.dim() != 2 or t.dtype != torch.bfloat16 or not t.is_contiguous() or t.shape != a2d.shape:
            raise B200Error("colsum_mul: need two contiguous bf16 [rows, cols] tensors of equal shape")
    rows, cols = a2d.shape
    if out is None:
        out = torch.empty(cols, dtype=torch.float32, device=a2d.device)
        accumulate = False
    rc = _lib.load().b200_colsum_mul(a2d.data_ptr(), b2d.data_ptr(), _ptr(row_scale), out.data_ptr(), rows, cols,
                                     int(rows_per_scale), int(accumulate), _stream())
    _lib.check(rc, "b200_colsum_mul")
    _count(1 if accumulate else 2)
    return out


# --------------------------------------------------------------------------------------------- per-op timing (tools)
_op_timing = None


def enable_op_timing(flag):
    """tools/bench_model.py --detail: bracket every op wrapper of this module with CUDA events (by op name)."""
    global _op_timing
    _op_timing = {} if flag else None


def op_timing_summary():
    if not _op_timing:
        return {}
    torch.cuda.synchronize()
    return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in _op_timing.items()}


def _timed(name, fn):
    def wrapper(*args, **kwargs):
        if _op_timing is None:
            return fn(*args, **kwargs)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn(*args, **kwargs)
        b.record()
        key = name
        if name == "gemm":
            key = f"gemm M={args[2]} N={args[3]} K={args[4]} epi={kwargs.get('epilogue', 0)}"
        _op_timing.setdefault(key, []).append((a, b))
        return out
    wrapper.__name__ = fn.__name__
    wrapper.__doc__ = fn.__doc__
    return wrapper


for _n in ("layernorm_fwd", "layernorm_bwd", "gemm", "attention_fwd", "attention_bwd", "patchify", "assemble_tokens",
           "colsum", "convert_f32", "dwconv_nhwc", "dwconv_nhwc_ln", "dwconv_nhwc_wgrad", "unpatchify_nhwc", "patchify_nhwc", "avgpool_nhwc", "naflex_assemble"):
    globals()[_n] = _timed(_n, globals()[_n])
