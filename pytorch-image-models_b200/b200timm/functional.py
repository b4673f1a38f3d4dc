"""Fused forward (inference) of timm's ViT hot path on the libb200timm kernels.

The functions are duck-typed on timm's own attribute names (``blk.norm1``, ``blk.attn.qkv``, ``blk.mlp.fc1``,
``model.patch_embed.proj`` ...) so the same code drives the mirror modules of b200timm.layers /
b200timm.vision_transformer and -- through b200timm.plugin -- timm's unmodified classes.

Fusion map per transformer block (reference: timm/models/vision_transformer.py:206-214):
    LN1            -> folded into the qkv GEMM (below); a layernorm kernel for the first block  (norm.py:85-90)
    qkv            -> GEMM + bias epilogue                          (attention.py:115)
    sdpa           -> tcgen05 attention, reads qkv in place         (attention.py:115-135)
    proj + ls1 + residual -> GEMM epilogue  x + gamma*(acc+b)       (attention.py:137, layer_scale.py:27, :212)
    LN2            -> folded into the fc1 GEMM
    fc1 + GELU     -> GEMM epilogue                                 (mlp.py:48-49)
    fc2 + ls2 + residual -> GEMM epilogue                           (mlp.py:52, :213)
Folded LayerNorm (inference; B200_LN_FOLD=0 switches it off): the residual GEMM that produces x also emits per-row partial
sums (sum x, sum x^2 of the bf16 values it stores); the Linear that follows norm(x) runs on x itself against the weight
pre-multiplied by the norm's gamma, and its epilogue finishes LN(x) W^T + b = rstd * (x W'^T - mean * rowsum(W')) + (W beta + b).
No LayerNorm pass over HBM remains between the blocks (b200_gemm_desc.ln_*).
Anything the kernels do not cover (qk-norm, attention dropout, non-GELU activations, head_dim != 64, ...)
raises B200Error: there is no PyTorch fallback on this path.
"""
import torch
from torch import nn
import torch.nn.functional as F

from . import ops
from ._lib import B200Error


def _is_identity(m):
    return m is None or isinstance(m, nn.Identity)


def _needs_grad(x, *modules):
    """True when this call must be recorded for backward: grad mode on and the input or any parameter of the
    given modules requires grad (training).  Inference (no_grad / inference_mode) takes the fused path."""
    if not torch.is_grad_enabled():
        return False
    if isinstance(x, torch.Tensor) and x.requires_grad:
        return True
    return any(p.requires_grad for m in modules if m is not None for p in m.parameters())


def _ln_params(norm):
    if _is_identity(norm):
        return None
    if type(norm).__name__ == "RmsNorm":     # this mirror's or timm's (timm/layers/norm.py:202-248)
        return norm.weight, None, ops.RmsEps(norm.eps)
    if not isinstance(norm, nn.LayerNorm):
        raise B200Error(f"unsupported norm layer {type(norm).__name__} (LayerNorm / RmsNorm only)")
    return norm.weight, norm.bias, norm.eps


# ------------------------------------------------------------------------------------------------ folded LayerNorm
import os as _os

_LN_FOLD = _os.environ.get("B200_LN_FOLD", "1")[:1] != "0"


def _ln_fold(norm, linear, w, b, tag):
    """Operands of `linear(norm(x))` evaluated on x itself (b200_gemm_desc.ln_*): (W * gamma as bf16 [N, K], c1 = row sums of
    that operand fp32 [N], c2 = W beta + b fp32 [N], eps), cached on the Linear and rebuilt when any of the four parameters
    changes; None when the pair is not foldable (RmsNorm / no affine / narrow output).  w, b: the GEMM operands the unfused
    path would use (bf16 cast / head-padded views of linear.weight / linear.bias)."""
    if not _LN_FOLD or type(norm) is not nn.LayerNorm and type(norm).__name__ != "LayerNorm":
        return None
    g, beta = norm.weight, norm.bias
    if g is None or beta is None or w.shape[0] <= 128:
        return None
    key = (ops.wkey(linear.weight), None if linear.bias is None else ops.wkey(linear.bias), ops.wkey(g), ops.wkey(beta),
           tuple(w.shape))

    def build():
        wf = w.detach().float()
        wg = (wf * g.detach().float()[None, :]).to(torch.bfloat16).contiguous()
        c2 = wf @ beta.detach().float()
        if b is not None:
            c2 = c2 + b.detach().float()
        return wg, wg.float().sum(1).contiguous(), c2.contiguous()
    return (*_cached(linear, "_b200_lnfold_" + tag, key, build), float(norm.eps))


class LnCarry:
    """Row statistics of the residual stream, handed from the GEMM that produced it (a block's fc2) to the next block's
    norm1 by the loop that calls the blocks (vit_forward_features, NaFlexVit.forward_features): `stats` describes exactly the
    tensor the previous block_forward returned.  Calls without a carrier (a block invoked on its own, timm's own loop through
    the plugin's per-module rebinding) fold norm2 only -- nothing is ever attached to tensors."""
    __slots__ = ("stats",)

    def __init__(self):
        self.stats = None


def _gamma(ls):
    if _is_identity(ls):
        return None
    g = getattr(ls, "gamma", None)
    if g is None:
        raise B200Error(f"unsupported layer-scale module {type(ls).__name__}")
    return g


def _check_drop(m, what):
    if _is_identity(m):
        return
    p = getattr(m, "p", getattr(m, "drop_prob", 0.0))
    if p and m.training:
        raise B200Error(f"{what} with p={p} in training mode is not supported by the fused inference path")


def _drop_path_factor(dp, x):
    """Per-sample DropPath factor (keep-mask / keep_prob) as fp32 [B], or None when the module is inactive.  The
    Bernoulli draw is made exactly as the reference makes it -- x.new_empty((B, 1, 1)).bernoulli_(keep_prob) on the
    torch CUDA generator (timm/layers/drop.py:170-175) -- so seeds, RNG consumption and gradient checkpointing
    replay match the reference; applying the factor is fused into the residual GEMM epilogue."""
    if _is_identity(dp):
        return None
    p = getattr(dp, "drop_prob", 0.0)
    if not p or not dp.training:
        return None
    keep = 1.0 - p
    mask = x.new_empty((x.shape[0], 1, 1)).bernoulli_(keep)
    f = mask.reshape(-1).float()
    if keep > 0.0 and getattr(dp, "scale_by_keep", True):
        f = f / keep
    return f.contiguous()


def check_attention(attn, allow_norms=False):
    """Raise unless `attn` is a timm Attention the kernels cover (attention.py:43-139).  The fused block paths take the
    plain module only; qk_norm / scale_norm run through attention_forward (allow_norms=True)."""
    if not allow_norms:
        if not _is_identity(getattr(attn, "q_norm", None)) or not _is_identity(getattr(attn, "k_norm", None)):
            raise B200Error("qk_norm attention is not supported")
        if not _is_identity(getattr(attn, "norm", None)):
            raise B200Error("scale_norm attention is not supported")
    head_pad(attn.head_dim)
    _check_drop(attn.attn_drop, "attn_drop")
    _check_drop(attn.proj_drop, "proj_drop")


def _attn_norms(attn):
    """((q_weight, q_bias, k_weight, k_bias, eps) or None, (weight, bias, eps) or None) of Attention's optional norms."""
    qn, kn, sn = getattr(attn, "q_norm", None), getattr(attn, "k_norm", None), getattr(attn, "norm", None)
    qk = None
    if not _is_identity(qn) or not _is_identity(kn):
        if _is_identity(qn) or _is_identity(kn) or type(qn) is not type(kn):
            raise B200Error("q_norm and k_norm must be the same kind of layer")
        (qw, qb, qe), (kw, kb, ke) = _ln_params(qn), _ln_params(kn)
        if float(qe) != float(ke) or qw is None or kw is None or (qb is None) != (kb is None):
            raise B200Error("q_norm / k_norm need equal eps and affine parameters")
        qk = (qw, qb, kw, kb, qe)
    sp = None if _is_identity(sn) else _ln_params(sn)
    if sp is not None and sp[0] is None:
        raise B200Error("scale_norm without affine parameters is not supported")
    return qk, sp


def mlp_act(mlp):
    """b200_gemm_desc.act code of an Mlp's activation: exact GELU (nn.GELU / timm GELU), tanh GELU (timm 'gelu_tanh' =
    GELUTanh, nn.GELU(approximate='tanh'); timm/layers/activations.py, create_act.py) or SiLU."""
    a = mlp.act
    name = type(a).__name__
    if name == "GELUTanh" or (name == "GELU" and getattr(a, "approximate", "none") == "tanh"):
        return ops.ACTS["gelu_tanh"]
    if name == "GELU" and getattr(a, "approximate", "none") == "none":
        return ops.ACTS["gelu"]
    if name in ("SiLU", "Swish"):
        return ops.ACTS["silu"]
    raise B200Error(f"unsupported Mlp activation {a} (GELU, tanh GELU, SiLU)")


def check_mlp(mlp, allow_norm=False):
    mlp_act(mlp)
    if not allow_norm and not _is_identity(getattr(mlp, "norm", None)):
        raise B200Error("Mlp with an inner norm runs through mlp_forward, not the fully fused block")
    if not isinstance(mlp.fc1, nn.Linear):
        raise B200Error("conv Mlp is not supported here")
    _check_drop(getattr(mlp, "drop1", None), "mlp drop")
    _check_drop(getattr(mlp, "drop2", None), "mlp drop")


def _bf16_2d(x, cols):
    if x.dtype != torch.bfloat16:
        raise B200Error(f"activations must be bf16 on the fused path, got {x.dtype}")
    x2 = x.reshape(-1, cols)
    return x2 if x2.is_contiguous() else x2.contiguous()


def _w(linear):
    """The bf16 GEMM operand of a Linear / Conv2d.  bf16 parameters (benchmark.py --precision bfloat16) are used as they
    are.  fp32 parameters (train.py --amp: fp32 master weights under torch.autocast) are cast the way autocast casts
    them: a differentiable `.to(bfloat16)` when the call is recorded for backward -- autograd widens the bf16 weight
    gradient back to fp32 exactly as it does for an autocast F.linear -- and a cached cast otherwise."""
    w = linear.weight
    if w.dtype == torch.bfloat16:
        return w if w.is_contiguous() else w.contiguous()
    if w.dtype != torch.float32:
        raise B200Error(f"weights must be bf16 or fp32, got {w.dtype}")
    if torch.is_grad_enabled() and w.requires_grad:
        return w.to(torch.bfloat16)
    key = ops.wkey(w)
    cache = getattr(linear, "_b200_wbf16", None)
    if cache is None or cache[0] != key:
        linear._b200_wbf16 = cache = (key, w.detach().to(torch.bfloat16).contiguous())
    return cache[1]


# ------------------------------------------------------------------------------------------------ head_dim != 64
# The attention kernels are built for head widths 64 (one 128-byte swizzle atom per token row and head) and 128 (two atoms:
# the long-sequence kernels with NH = 2).  Other widths (attention.py:85-93 allows any dim // num_heads or attn_head_dim: 32 in
# timm's test_vit* models, 72 in SigLIP SO400M, 80 in ViT-H, 96) run on the SAME kernels through zero
# padding that lives entirely in the WEIGHTS: the rows of the qkv projection (and its bias) are spread to P = 64 / 128 per head
# with zero rows in between, the columns of the output projection likewise -- the GEMM then writes a [M, 3 * H * P] buffer whose
# padding columns are exact zeros (q . k over the padding adds 0, the padded v columns produce output columns that meet zero
# weights), and softmax scaling uses the module's own scale (true head_dim ** -0.5).  Padded weights are cached at inference
# and built differentiably (F.pad on a view) in training, so gradients reach the real parameters through autograd.
def head_pad(hd):
    """Padded head width P the kernels run (64 or 128), or raises for widths that cannot be padded."""
    if hd % 8 == 0 and 0 < hd <= 64:
        return 64
    if hd % 8 == 0 and 64 < hd <= 128:
        return 128
    raise B200Error(f"head_dim {hd} is not supported (a multiple of 8 up to 128; 64 and 128 natively, the others through "
                    f"weight padding)")


def _cached(module, attr, key, build):
    cache = getattr(module, attr, None)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            cache = (key, build())
        setattr(module, attr, cache)
    return cache[1]


def _pad_rows(t, groups, H, hd):
    """[groups * H * hd, ...] -> [groups * H * P, ...]: every head's hd rows followed by P - hd zero rows (P = head_pad(hd))."""
    P = head_pad(hd)
    if t is None or hd == P:
        return t
    tail = t.shape[1:]
    v = t.reshape(groups * H, hd, *tail)
    pad = [0, 0] * len(tail) + [0, P - hd]
    return F.pad(v, pad).reshape(groups * H * P, *tail)


def _pad_cols(w, H, hd):
    """[out, H * hd] -> [out, H * P]: zero columns after every head's hd columns."""
    P = head_pad(hd)
    if hd == P:
        return w
    return F.pad(w.reshape(w.shape[0], H, hd), (0, P - hd)).reshape(w.shape[0], H * P)


def head_padded_linear_in(linear, groups, H, hd, bias=None, weight=None):
    """(weight, bias) of a projection INTO head space (qkv: groups = 3, kv: 2, q: 1) with heads padded to head_pad(hd) rows."""
    w = _w(linear) if weight is None else weight
    b = linear.bias if bias is None else bias
    if hd == head_pad(hd):
        return w, b
    if torch.is_grad_enabled() and (w.requires_grad or (b is not None and b.requires_grad)):
        return _pad_rows(w, groups, H, hd), _pad_rows(b, groups, H, hd)
    key = (ops.wkey(w), None if b is None else ops.wkey(b), groups, H, hd)
    return _cached(linear, "_b200_hpad_in", key, lambda: (_pad_rows(w.detach(), groups, H, hd).contiguous(),
                                                         None if b is None else _pad_rows(b.detach(), groups, H, hd).contiguous()))


def head_padded_linear_out(linear, H, hd):
    """weight of a projection OUT of head space (attention's proj) with every head's columns padded to head_pad(hd)."""
    w = _w(linear)
    if hd == head_pad(hd):
        return w
    if torch.is_grad_enabled() and w.requires_grad:
        return _pad_cols(w, H, hd)
    return _cached(linear, "_b200_hpad_out", (ops.wkey(w), H, hd), lambda: _pad_cols(w.detach(), H, hd).contiguous())


def _padded_out_linear(linear, x2):
    """F.linear whose out_features is not a multiple of 8 (classifier heads with e.g. 10 classes): the GEMM
    runs on a zero-padded copy of weight / bias (cached on the module, refreshed when the parameters change)
    and the padding columns are sliced off."""
    w, b = linear.weight, linear.bias
    N, K = w.shape
    npad = (N + 7) // 8 * 8
    key = (*ops.wkey(w), None if b is None else ops.wkey(b))
    cache = getattr(linear, "_b200_npad", None)
    if cache is None or cache[0] != key:
        wp = torch.zeros(npad, K, dtype=torch.bfloat16, device=w.device)
        wp[:N] = w.detach()
        bp = None
        if b is not None:
            bp = torch.zeros(npad, dtype=b.dtype, device=b.device)
            bp[:N] = b.detach()
        cache = (key, wp, bp)
        linear._b200_npad = cache
    return ops.linear_fwd(x2, cache[1], cache[2])[:, :N]


def layernorm(norm, x):
    """LayerNorm.forward (timm/layers/norm.py:85-90)."""
    w, b, eps = _ln_params(norm)
    if _needs_grad(x, norm):
        from .autograd import LayerNormFn
        return LayerNormFn.apply(x, w, b, eps)
    return ops.layernorm_fwd(x, w, b, eps)[0]


def attention_forward(attn, x, kv_len=None, residual=None, gamma=None):
    """Attention.forward (timm/layers/attention.py:108-139) incl. qk_norm / scale_norm.  With `residual` (inference only)
    the block's `x + ls(attn(...))` is folded into the output projection's epilogue."""
    check_attention(attn, allow_norms=True)
    qk, sn = _attn_norms(attn)
    B, N, C = x.shape
    H, hd = attn.num_heads, attn.head_dim
    if hd != 64 and (qk is not None or sn is not None):
        raise B200Error("qk_norm / scale_norm with head_dim != 64 is not covered")
    wqkv, bqkv = head_padded_linear_in(attn.qkv, 3, H, hd)
    wproj = head_padded_linear_out(attn.proj, H, hd)
    if _needs_grad(x, attn):
        if residual is not None:
            raise B200Error("attention_forward: the fused residual is an inference-only epilogue")
        _bf16_2d(x, C)
        if qk is None and sn is None:
            from .autograd import AttentionFn
            return AttentionFn.apply(x, wqkv, bqkv, wproj, attn.proj.bias, kv_len, H, attn.scale)
        from .autograd import AttentionNormFn
        qw, qb, kw, kb, qe = qk if qk is not None else (None, None, None, None, 0.0)
        nw, nb, ne = sn if sn is not None else (None, None, 0.0)
        return AttentionNormFn.apply(x, wqkv, bqkv, wproj, attn.proj.bias, qw, qb, kw, kb, nw, nb, kv_len, H, attn.scale, qe, ne)
    x2 = _bf16_2d(x, C)
    qkv = ops.linear_fwd(x2, wqkv, bqkv)
    if qk is not None:
        ops.qk_norm_fwd(qkv, B * N, H, *qk)
    o, _ = ops.attention_fwd(qkv, B, N, H, head_pad(hd), attn.scale, kv_len)
    o2 = o.view(B * N, -1)
    if sn is not None:
        o2 = ops.layernorm_fwd(o2, *sn)[0]
    if residual is not None:
        y = ops.linear_fwd(o2, wproj, attn.proj.bias, epilogue=ops.EPI_BIAS_RESIDUAL, aux=residual, gamma=gamma)
    else:
        y = ops.linear_fwd(o2, wproj, attn.proj.bias)
    return y.view(B, N, -1)


def _rope_table(rope, n_rot, head_dim, device, batch=None):
    """The `rope` argument of AttentionRope.forward -> contiguous fp32 [n_rot, 2 * head_dim] = [sin | cos] on `device`, or
    [B, n_rot, 2 * head_dim] for a per-sample table (the reference broadcasts `emb` against q[B, H, n_rot, hd]: shapes
    [n_rot, 2 hd], [1, 1, n_rot, 2 hd] or [B, 1, n_rot, 2 hd], pos_embed_sincos.py:281-297)."""
    if head_dim != 64:
        raise B200Error("AttentionRope: only head_dim 64 is covered")
    r = rope
    if r.dim() == 4 and r.shape[1] == 1:
        r = r[:, 0]
    while r.dim() > 2 and r.shape[0] == 1:
        r = r[0]
    if r.dim() == 3 and batch is not None and tuple(r.shape) == (batch, n_rot, 2 * head_dim):
        pass
    elif r.dim() != 2 or tuple(r.shape) != (n_rot, 2 * head_dim):
        raise B200Error(f"AttentionRope: rope must broadcast from [{n_rot}, {2 * head_dim}] (or [B, 1, {n_rot}, {2 * head_dim}]), "
                        f"got {tuple(rope.shape)}")
    key = (*ops.wkey(rope), str(device), tuple(r.shape))
    cache = _rope_table.cache
    if cache.get("key") != key:
        cache["key"], cache["val"] = key, r.detach().to(device=device, dtype=torch.float32).contiguous()
    return cache["val"]


_rope_table.cache = {}


def _qkv_params(attn):
    """(weight [3A, C] bf16, bias [3A] or None) of the q / k / v projection(s) of AttentionRope / EvaAttention: the fused `qkv`
    Linear (EvaAttention keeps q_bias / v_bias parameters and a zero k_bias buffer, eva.py:157-165, 209-217), or -- for
    `qkv_fused=False` -- q_proj / k_proj / v_proj stacked into one [3A, C] operand (a differentiable cat of the WEIGHTS in
    training, cached otherwise), so that one GEMM produces the packed [M, 3A] buffer the attention kernels read."""
    if attn.qkv is not None:
        if getattr(attn, "q_bias", None) is not None:
            return _w(attn.qkv), torch.cat((attn.q_bias, attn.k_bias, attn.v_bias))
        return _w(attn.qkv), attn.qkv.bias
    lins = (attn.q_proj, attn.k_proj, attn.v_proj)
    ws = [_w(l) for l in lins]
    bs = [l.bias for l in lins]
    has_bias = any(b is not None for b in bs)

    def cat_bias():
        ref = next(b for b in bs if b is not None)
        return torch.cat([b if b is not None else torch.zeros(l.weight.shape[0], dtype=ref.dtype, device=ref.device)
                          for b, l in zip(bs, lins)])
    if torch.is_grad_enabled() and any(w.requires_grad for w in ws):
        return torch.cat(ws, 0), (cat_bias() if has_bias else None)
    key = tuple(ops.wkey(w) for w in ws) + tuple(None if b is None else ops.wkey(b) for b in bs)
    return _cached(attn, "_b200_qkvcat", key, lambda: (torch.cat([w.detach() for w in ws], 0).contiguous(),
                                                       cat_bias().detach() if has_bias else None))


def attention_rope_forward(attn, x, rope=None, kv_len=None):
    """AttentionRope.forward (timm/layers/attention.py:224-282) and EvaAttention.forward (timm/models/eva.py:197-262): QKV
    projection (fused or separate q / k / v), optional q / k normalisation, rotary embedding of q / k (all tokens after the
    first num_prefix_tokens; one table for the batch or one per sample), attention, optional scale-norm, output projection.
    The plain configuration (fused qkv, no norms, shared interleaved table) rotates in the QKV GEMM's epilogue."""
    if attn.training and (attn.attn_drop.p > 0 or attn.proj_drop.p > 0):
        raise B200Error("AttentionRope: dropout is not covered")
    B, N, C = x.shape
    H, hd = attn.num_heads, attn.head_dim
    if hd != 64:
        raise B200Error("AttentionRope: only head_dim 64 is covered")
    half = bool(getattr(attn, "rotate_half", False))
    npt = attn.num_prefix_tokens
    emb = None if rope is None else _rope_table(rope, N - npt, hd, x.device, batch=B)
    qk, sn = _attn_norms(attn)
    wqkv, bqkv = _qkv_params(attn)
    plain = qk is None and sn is None and (emb is None or emb.dim() == 2)
    if _needs_grad(x, attn):
        _bf16_2d(x, C)
        if plain:
            from .autograd import AttentionRopeFn
            return AttentionRopeFn.apply(x, wqkv, bqkv, _w(attn.proj), attn.proj.bias, emb, kv_len, H, attn.scale, npt, half)
        from .autograd import AttentionNormFn
        qw, qb, kw, kb, qe = qk if qk is not None else (None, None, None, None, 0.0)
        nw, nb, ne = sn if sn is not None else (None, None, 0.0)
        return AttentionNormFn.apply(x, wqkv, bqkv, _w(attn.proj), attn.proj.bias, qw, qb, kw, kb, nw, nb, kv_len, H, attn.scale,
                                     qe, ne, emb, npt, half)
    x2 = _bf16_2d(x, C)
    if plain:
        qkv = rope_qkv_projection(x2, wqkv, bqkv, emb, B, N, H, npt, half)
    else:
        qkv = ops.linear_fwd(x2, wqkv, bqkv)
        if qk is not None:
            ops.qk_norm_fwd(qkv, B * N, H, *qk)
        if emb is not None:
            ops.rope_qk(qkv, emb, B, N, H, npt, half=half)
    o, _ = ops.attention_fwd(qkv, B, N, H, hd, attn.scale, kv_len)
    o2 = o.view(B * N, -1)
    if sn is not None:
        o2 = ops.layernorm_fwd(o2, *sn)[0]
    y = ops.linear_fwd(o2, _w(attn.proj), attn.proj.bias)
    return y.view(B, N, -1)


def rope_qkv_projection(x2, wqkv, bqkv, emb, B, N, H, npt, half):
    """qkv = x W^T + b with q / k rotated: in the GEMM epilogue for the interleaved layout (neighbouring columns pair
    up inside one thread's registers), by one in-place pass for the half layout (partners are 32 columns apart)."""
    if emb is None:
        return ops.linear_fwd(x2, wqkv, bqkv)
    if not half:
        return ops.linear_fwd(x2, wqkv, bqkv, epilogue=ops.EPI_BIAS_ROPE, rope=(emb, N, npt, 2 * H * 64))
    return ops.rope_qk(ops.linear_fwd(x2, wqkv, bqkv), emb, B, N, H, npt, half=True)


def mlp_forward(mlp, x, residual=None, gamma=None, row_scale=None, rows_per_scale=0):
    """Mlp.forward (timm/layers/mlp.py:47-54): fc2(norm(GELU(fc1(x)))), GELU fused into fc1's epilogue, the optional inner
    norm (Mlp(norm_layer=...): `scale_mlp_norm` blocks) as a LayerNorm kernel between the two GEMMs.  With `residual`
    (inference only) the block's `x + drop_path(ls(mlp(...)))` is folded into fc2's epilogue."""
    check_mlp(mlp, allow_norm=True)
    shape = x.shape
    norm = None if _is_identity(getattr(mlp, "norm", None)) else _ln_params(mlp.norm)
    if _needs_grad(x, mlp):
        from .autograd import MlpFn
        if residual is not None:
            raise B200Error("mlp_forward: the fused residual is an inference-only epilogue")
        _bf16_2d(x, shape[-1])
        nw, nb, eps = norm if norm is not None else (None, None, 0.0)
        return MlpFn.apply(x, _w(mlp.fc1), mlp.fc1.bias, _w(mlp.fc2), mlp.fc2.bias, mlp_act(mlp), nw, nb, eps)
    x2 = _bf16_2d(x, shape[-1])
    h = ops.linear_fwd(x2, _w(mlp.fc1), mlp.fc1.bias, epilogue=ops.EPI_BIAS_GELU, act=mlp_act(mlp))
    if norm is not None:
        h = ops.layernorm_fwd(h, *norm)[0]
    if residual is not None:
        y = ops.linear_fwd(h, _w(mlp.fc2), mlp.fc2.bias, epilogue=ops.EPI_BIAS_RESIDUAL, aux=residual, gamma=gamma,
                           row_scale=row_scale, rows_per_scale=rows_per_scale)
    else:
        y = ops.linear_fwd(h, _w(mlp.fc2), mlp.fc2.bias)
    return y.view(*shape[:-1], -1)


def is_glu(mlp):
    """GluMlp / SwiGLUPacked (one fc1 of width 2H) or SwiGLU (fc1_g + fc1_x), timm/layers/mlp.py:57-165."""
    return hasattr(mlp, "fc1_g") or hasattr(mlp, "gate_last")


def _glu_params(mlp):
    """(w1 [2H, D] bf16, b1, gate_first, act name, norm params or None) of a gated MLP; the gate rows come first for SwiGLU
    (fc1_g stacked on fc1_x -- a differentiable cat when training, cached otherwise)."""
    act = {"SiLU": "silu", "Sigmoid": "sigmoid"}.get(type(mlp.act).__name__)
    if act is None:
        raise B200Error(f"unsupported gate activation {mlp.act} (SiLU / sigmoid)")
    _check_drop(getattr(mlp, "drop1", None), "mlp drop")
    _check_drop(getattr(mlp, "drop2", None), "mlp drop")
    if hasattr(mlp, "fc1_g"):
        wg, wx = _w(mlp.fc1_g), _w(mlp.fc1_x)
        bg, bx = mlp.fc1_g.bias, mlp.fc1_x.bias
        if torch.is_grad_enabled() and (wg.requires_grad or wx.requires_grad):
            w1 = torch.cat([wg, wx], 0)
            b1 = None if bg is None else torch.cat([bg, bx], 0)
        else:
            key = (*ops.wkey(wg), *ops.wkey(wx),
                   None if bg is None else (*ops.wkey(bg), *ops.wkey(bx)))
            cache = getattr(mlp, "_b200_w1cat", None)
            if cache is None or cache[0] != key:
                cache = (key, torch.cat([wg, wx], 0).contiguous(), None if bg is None else torch.cat([bg, bx], 0))
                mlp._b200_w1cat = cache
            w1, b1 = cache[1], cache[2]
        gate_first = True
    else:
        if not isinstance(mlp.fc1, nn.Linear):
            raise B200Error("conv GluMlp is not supported")
        w1, b1, gate_first = _w(mlp.fc1), mlp.fc1.bias, not mlp.gate_last
    if w1.shape[0] % 16:
        raise B200Error(f"gated MLP hidden width {w1.shape[0] // 2} is not a multiple of 8")
    norm = None if _is_identity(getattr(mlp, "norm", None)) else _ln_params(mlp.norm)
    if norm is not None and isinstance(norm[2], ops.RmsEps):
        raise B200Error("gated MLP with an inner RMSNorm is not supported")
    return w1, b1, gate_first, act, norm


def glu_mlp_forward(mlp, x, residual=None, gamma=None, row_scale=None, rows_per_scale=0):
    """GluMlp.forward / SwiGLU.forward (timm/layers/mlp.py:104-113, 153-162).  With `residual` (inference only) the block's
    `x + drop_path(ls(mlp(...)))` is folded into fc2's epilogue."""
    w1, b1, gate_first, act, norm = _glu_params(mlp)
    shape = x.shape
    if _needs_grad(x, mlp):
        from .autograd import GluMlpFn
        if residual is not None:
            raise B200Error("glu_mlp_forward: the fused residual is an inference-only epilogue")
        _bf16_2d(x, shape[-1])
        nw, nb, eps = norm if norm is not None else (None, None, 0.0)
        return GluMlpFn.apply(x, w1, b1, _w(mlp.fc2), mlp.fc2.bias, nw, nb, eps, gate_first, act)
    x2 = _bf16_2d(x, shape[-1])
    a = ops.glu_fwd(ops.linear_fwd(x2, w1, b1), gate_first, act)
    if norm is not None:
        a = ops.layernorm_fwd(a, *norm)[0]
    if residual is not None:
        y = ops.linear_fwd(a, _w(mlp.fc2), mlp.fc2.bias, epilogue=ops.EPI_BIAS_RESIDUAL, aux=residual, gamma=gamma,
                           row_scale=row_scale, rows_per_scale=rows_per_scale)
    else:
        y = ops.linear_fwd(a, _w(mlp.fc2), mlp.fc2.bias)
    return y.view(*shape[:-1], -1)


def _variant_block_forward(blk, x, kv_len):
    """Block.forward (vision_transformer.py:206-214) for blocks outside the fully fused BlockFn: a gated mlp and / or an
    attention with qk_norm / scale_norm.  Inference: every residual (+ LayerScale) rides in the epilogue of the branch's
    last GEMM.  Training: the per-module autograd functions composed the way the reference composes its modules
    (residual adds / LayerScale / DropPath as tensor ops)."""
    attn, mlp = blk.attn, blk.mlp
    B, N, D = x.shape
    run_mlp = glu_mlp_forward if is_glu(mlp) else mlp_forward
    if _needs_grad(x, blk):
        def branch(y, ls, dp):
            g = _gamma(ls)
            if g is not None:
                y = y * g.to(y.dtype)
            f = _drop_path_factor(dp, y)
            return y if f is None else y * f.view(B, 1, 1).to(y.dtype)
        x = x + branch(attention_forward(attn, layernorm(blk.norm1, x), kv_len), blk.ls1, blk.drop_path1)
        return x + branch(run_mlp(mlp, layernorm(blk.norm2, x)), blk.ls2, blk.drop_path2)
    x2 = _bf16_2d(x, D)
    h = ops.layernorm_fwd(x2, *_ln_params(blk.norm1))[0]
    x2 = attention_forward(attn, h.view(B, N, D), kv_len, residual=x2, gamma=_gamma(blk.ls1)).view(B * N, D)
    h = ops.layernorm_fwd(x2, *_ln_params(blk.norm2))[0]
    if is_glu(mlp):
        return glu_mlp_forward(mlp, h, residual=x2, gamma=_gamma(blk.ls2)).view(B, N, D)
    return mlp_forward(mlp, h, residual=x2, gamma=_gamma(blk.ls2)).view(B, N, D)


def _act_code(act_module):
    name = type(act_module).__name__
    if name == "GELUTanh" or (name == "GELU" and getattr(act_module, "approximate", "none") == "tanh"):
        return ops.ACTS["gelu_tanh"]
    if name == "GELU" and getattr(act_module, "approximate", "none") == "none":
        return ops.ACTS["gelu"]
    if name in ("SiLU", "Swish"):
        return ops.ACTS["silu"]
    raise B200Error(f"unsupported activation {act_module} (GELU, tanh GELU, SiLU)")


def is_parallel_block(blk):
    """timm's ParallelScalingBlock (vision_transformer.py:292-420): attention and MLP read ONE normalised input through ONE
    fused in-projection and are summed before the residual (ViT-22B; vit_*_xp_224, test_vit3)."""
    return hasattr(blk, "in_proj") and hasattr(blk, "in_split") and hasattr(blk, "in_norm")


def parallel_block_forward(blk, x, kv_len=None):
    """ParallelScalingBlock.forward (vision_transformer.py:373-420):

        y = in_norm(x);  [x_mlp | q | k | v] = in_proj(y) (+ mlp_bias);  q, k = q_norm(q), k_norm(k)
        out = x + drop_path(ls(out_proj(cat(sdpa(q, k, v), act(x_mlp)))))          (or attn_out_proj(.) + mlp_out_proj(.))

    The fused in-projection runs as two GEMMs over row ranges of in_proj.weight (views, no copy): the MLP rows with the
    activation in the epilogue, the q / k / v rows into the contiguous [M, 3D] buffer the attention kernels read in place.
    The output projection runs as two GEMMs over column ranges of out_proj.weight (strided views) -- the concatenation is
    never materialised; at inference both carry the residual epilogue (x + gamma * rs * (...)), so LayerScale, DropPath
    and the residual add cost nothing.  Training composes AttentionFn / AttentionNormFn and MlpFn on those views."""
    B, N, D = x.shape
    H, hd = blk.num_heads, blk.head_dim
    head_pad(hd)
    _check_drop(blk.attn_drop, "attn_drop")
    _check_drop(blk.mlp_drop, "mlp drop")
    act = _act_code(blk.mlp_act)
    hidden = blk.in_split[0]
    if list(blk.in_split[1:]) != [D, D, D] or hidden % 8 or D % 8:
        raise B200Error("ParallelScalingBlock: unexpected in_split / widths that are not multiples of 8")
    qk = None
    if not _is_identity(blk.q_norm) or not _is_identity(blk.k_norm):
        (qw, qb, qe), (kw, kb, ke) = _ln_params(blk.q_norm), _ln_params(blk.k_norm)
        if float(qe) != float(ke) or type(blk.q_norm) is not type(blk.k_norm) or (qb is None) != (kb is None):
            raise B200Error("q_norm / k_norm need equal eps and affine parameters")
        qk = (qw, qb, kw, kb, qe)
    if hd != 64 and qk is not None:
        raise B200Error("qk_norm with head_dim != 64 is not covered")
    w_in = _w(blk.in_proj)
    w_mlp, w_qkv = w_in[:hidden], w_in[hidden:]
    b_in = blk.in_proj.bias
    b_mlp = blk.mlp_bias if b_in is None else b_in[:hidden]
    b_qkv = None if b_in is None else b_in[hidden:]
    if blk.out_proj is not None:
        w_out = _w(blk.out_proj)
        w_ao, w_mo = w_out[:, :D], w_out[:, D:]
        b_ao, b_mo = blk.out_proj.bias, None
    else:
        w_ao, w_mo = _w(blk.attn_out_proj), _w(blk.mlp_out_proj)
        b_ao, b_mo = blk.attn_out_proj.bias, blk.mlp_out_proj.bias
    if hd != head_pad(hd):      # heads padded inside the weights (see head_pad): not cached here -- slices of a fused projection
        w_qkv, b_qkv, w_ao = _pad_rows(w_qkv, 3, H, hd), _pad_rows(b_qkv, 3, H, hd), _pad_cols(w_ao, H, hd)
        if not (torch.is_grad_enabled() and w_qkv.requires_grad):
            w_qkv, w_ao = w_qkv.contiguous(), w_ao.contiguous()
    gamma = _gamma(blk.ls)
    if _needs_grad(x, blk):
        from .autograd import AttentionFn, AttentionNormFn, MlpFn
        _bf16_2d(x, D)
        y = layernorm(blk.in_norm, x)
        if qk is None:
            ya = AttentionFn.apply(y, w_qkv, b_qkv, w_ao, b_ao, kv_len, H, blk.scale)
        else:
            ya = AttentionNormFn.apply(y, w_qkv, b_qkv, w_ao, b_ao, qk[0], qk[1], qk[2], qk[3], None, None, kv_len, H, blk.scale,
                                       qk[4], 0.0)
        ym = MlpFn.apply(y, w_mlp, b_mlp, w_mo, b_mo, act)
        br = ya + ym
        if gamma is not None:
            br = br * gamma.to(br.dtype)
        f = _drop_path_factor(blk.drop_path, br)
        if f is not None:
            br = br * f.view(B, 1, 1).to(br.dtype)
        return x + br
    rs = _drop_path_factor(blk.drop_path, x)          # None in eval mode
    x2 = _bf16_2d(x, D)
    y = ops.layernorm_fwd(x2, *_ln_params(blk.in_norm))[0]
    hmlp = ops.linear_fwd(y, w_mlp, b_mlp, epilogue=ops.EPI_BIAS_GELU, act=act)
    qkv = ops.linear_fwd(y, w_qkv, b_qkv)
    if qk is not None:
        ops.qk_norm_fwd(qkv, B * N, H, *qk)
    o, _ = ops.attention_fwd(qkv, B, N, H, head_pad(hd), blk.scale, kv_len)
    t = ops.linear_fwd(o.view(B * N, -1), w_ao, b_ao, epilogue=ops.EPI_BIAS_RESIDUAL, aux=x2, gamma=gamma, row_scale=rs,
                       rows_per_scale=N)
    out = ops.linear_fwd(hmlp, w_mo, b_mo, epilogue=ops.EPI_BIAS_RESIDUAL, aux=t, gamma=gamma, row_scale=rs, rows_per_scale=N)
    return out.view(B, N, D)


def block_forward(blk, x, kv_len=None, carry=None):
    """Block.forward (timm/models/vision_transformer.py:206-214), eval mode.  carry: the calling loop's LnCarry (folded
    LayerNorm across blocks) or None."""
    st_in = None
    if carry is not None:
        st_in, carry.stats = carry.stats, None        # consumed here; set again below only by the fused inference path
    if is_parallel_block(blk):
        return parallel_block_forward(blk, x, kv_len)
    attn, mlp = blk.attn, blk.mlp
    qk, sn = _attn_norms(attn)
    if is_glu(mlp) or qk is not None or sn is not None or not _is_identity(getattr(mlp, "norm", None)):
        check_attention(attn, allow_norms=True)
        return _variant_block_forward(blk, x, kv_len)
    check_attention(attn)
    check_mlp(mlp)
    B, N, D = x.shape
    if _needs_grad(x, blk):
        from .autograd import BlockFn
        _bf16_2d(x, D)
        n1w, n1b, e1 = _ln_params(blk.norm1)
        n2w, n2b, e2 = _ln_params(blk.norm2)
        rs1, rs2 = _drop_path_factor(blk.drop_path1, x), _drop_path_factor(blk.drop_path2, x)
        wqkv, bqkv = head_padded_linear_in(attn.qkv, 3, attn.num_heads, attn.head_dim)
        return BlockFn.apply(x, n1w, n1b, wqkv, bqkv, head_padded_linear_out(attn.proj, attn.num_heads, attn.head_dim),
                             attn.proj.bias, n2w, n2b,
                             _w(mlp.fc1), mlp.fc1.bias, _w(mlp.fc2), mlp.fc2.bias, _gamma(blk.ls1), _gamma(blk.ls2),
                             rs1, rs2, kv_len, attn.num_heads, attn.scale, e1, e2, mlp_act(mlp))
    rs1, rs2 = _drop_path_factor(blk.drop_path1, x), _drop_path_factor(blk.drop_path2, x)   # None in eval mode
    M = B * N
    fold = _LN_FOLD and M >= 256 and D > 128       # the shapes the folded-LayerNorm GEMM instantiations serve
    if not fold or not x.is_contiguous():
        st_in = None                                # else: left by the GEMM that produced x (the previous block's fc2)
    x2 = _bf16_2d(x, D)
    wqkv, bqkv = head_padded_linear_in(attn.qkv, 3, attn.num_heads, attn.head_dim)
    f1 = _ln_fold(blk.norm1, attn.qkv, wqkv, bqkv, "qkv") if st_in is not None else None
    if f1 is not None:
        qkv = ops.linear_fwd(x2, f1[0], f1[2], ln_in=(st_in, f1[1], f1[3]))
    else:
        w1, b1, eps1 = _ln_params(blk.norm1)
        qkv = ops.linear_fwd(ops.layernorm_fwd(x2, w1, b1, eps1)[0], wqkv, bqkv)
    o, _ = ops.attention_fwd(qkv, B, N, attn.num_heads, head_pad(attn.head_dim), attn.scale, kv_len)
    act = mlp_act(mlp)
    f2 = _ln_fold(blk.norm2, mlp.fc1, _w(mlp.fc1), mlp.fc1.bias, "fc1") if fold else None
    st2 = ops.ln_stats_buffer(M, D, x.device) if f2 is not None else None
    x2 = ops.linear_fwd(o.view(B * N, -1), head_padded_linear_out(attn.proj, attn.num_heads, attn.head_dim), attn.proj.bias,
                        epilogue=ops.EPI_BIAS_RESIDUAL, aux=x2,
                        gamma=_gamma(blk.ls1), row_scale=rs1, rows_per_scale=N, ln_stats_out=st2)
    if f2 is not None:
        h = ops.linear_fwd(x2, f2[0], f2[2], epilogue=ops.EPI_BIAS_GELU, act=act, ln_in=(st2, f2[1], f2[3]))
    else:
        w2, b2, eps2 = _ln_params(blk.norm2)
        h = ops.layernorm_fwd(x2, w2, b2, eps2)[0]
        h = ops.linear_fwd(h, _w(mlp.fc1), mlp.fc1.bias, epilogue=ops.EPI_BIAS_GELU, act=act)
    # the statistics of the block's output ride along for the next block's norm1 when the calling loop carries them
    st3 = ops.ln_stats_buffer(M, D, x.device) if fold and carry is not None else None
    x2 = ops.linear_fwd(h, _w(mlp.fc2), mlp.fc2.bias, epilogue=ops.EPI_BIAS_RESIDUAL, aux=x2, gamma=_gamma(blk.ls2),
                        row_scale=rs2, rows_per_scale=N, ln_stats_out=st3)
    if carry is not None:
        carry.stats = st3
    return x2.view(B, N, D)


def patch_embed_forward(pe, img):
    """PatchEmbed.forward (timm/layers/patch_embed.py:117-142), flatten=True output [B, G, D]."""
    if not _is_identity(pe.norm):
        raise B200Error("PatchEmbed with a norm layer is not supported")
    ph, pw = pe.patch_size
    if ph != pw:
        raise B200Error("non-square patches are not supported")
    if getattr(pe, "dynamic_img_pad", False):      # patch_embed.py:131-134: zero-pad bottom / right up to a patch multiple
        pad_h = (ph - img.shape[-2] % ph) % ph
        pad_w = (pw - img.shape[-1] % pw) % pw
        if pad_h or pad_w:
            if img.dtype == torch.uint8:
                raise B200Error("dynamic_img_pad with raw uint8 batches is not covered")
            img = F.pad(img, (0, pad_w, 0, pad_h))
    B, C, H, W = img.shape
    if pe.img_size is not None and getattr(pe, "strict_img_size", True) and not getattr(pe, "dynamic_img_pad", False) \
            and (H, W) != tuple(pe.img_size):
        raise AssertionError(f"Input size ({H}x{W}) doesn't match model ({pe.img_size[0]}x{pe.img_size[1]}).")
    if H % ph or W % pw:
        raise AssertionError(f"Input size ({H}x{W}) should be divisible by patch size ({ph}).")
    w = _w(pe.proj)
    if _needs_grad(None, pe):
        from .autograd import PatchEmbedFn
        if img.dtype not in (torch.bfloat16, torch.float32, torch.uint8):
            raise B200Error("images must be bf16, fp32 or (with ops.set_input_norm) uint8 on the fused path")
        return PatchEmbedFn.apply(img, w, pe.proj.bias, ph)
    D = w.shape[0]
    K = C * ph * pw
    kpad = (K + 7) // 8 * 8
    w2 = w.view(D, K)
    if kpad != K:  # 16-byte row stride for TMA (e.g. patch 14: K = 588 -> 592), zero padded
        cache = getattr(pe, "_b200_wpad", None)
        if cache is None or cache[0] != ops.wkey(w):
            wp = torch.zeros(D, kpad, dtype=w.dtype, device=w.device)
            wp[:, :K] = w2
            pe._b200_wpad = (ops.wkey(w), wp)
            cache = pe._b200_wpad
        w2 = cache[1]
    if kpad == K and ops.patch_embed_supported(img, ph, D) and _IM2COL_FREE:
        # the image IS the GEMM operand (5-D tensor map): no patch matrix in HBM
        return ops.patch_embed_fwd(img, w2, pe.proj.bias, None, 0, (H // ph) * (W // pw), 0, ph)
    patches = ops.patchify(img.contiguous(), ph, kpad)
    y = ops.gemm(patches, w2, patches.shape[0], D, kpad, bias=pe.proj.bias)
    return y.view(B, (H // ph) * (W // pw), D)


import os as _os
_IM2COL_FREE = _os.environ.get("B200_IM2COL_FREE", "1") == "1"       # 1 = the image itself is the patch GEMM's A operand (0: gather path)


def _fused_patch_tokens(model, img, pos, prefix, n_prefix, pos_has_prefix):
    """patch_embed + _pos_embed of an inference forward in ONE GEMM launch (+ a few prefix rows): the image is the A operand,
    the position rows are the residual tile of the epilogue, the result lands behind the prefix tokens of [B, P + G, D].
    Returns None when the configuration needs the gather path."""
    pe = model.patch_embed
    ph, pw = pe.patch_size
    w = _w(pe.proj)
    D = w.shape[0]
    if (not _IM2COL_FREE or ph != pw or not _is_identity(pe.norm) or getattr(pe, "dynamic_img_pad", False)
            or not ops.patch_embed_supported(img, ph, D) or (pos is not None and pos.dtype != torch.bfloat16)
            or (prefix is not None and prefix.dtype != torch.bfloat16)):
        return None
    B, C, H, W = img.shape
    if pe.img_size is not None and getattr(pe, "strict_img_size", True) and (H, W) != tuple(pe.img_size):
        return None
    G = (H // ph) * (W // pw)
    pos2 = None if pos is None else pos.reshape(-1, D)
    if pos2 is not None and pos2.shape[0] != (n_prefix + G if pos_has_prefix else G):
        return None
    pos2 = None if pos2 is None else (pos2 if pos2.is_contiguous() else pos2.contiguous())
    x = ops.patch_embed_fwd(img, w.reshape(D, -1), pe.proj.bias, pos2, n_prefix if (pos_has_prefix and pos2 is not None) else 0,
                            n_prefix + G, n_prefix, ph)
    if n_prefix:
        x[:, :n_prefix] = prefix if (pos2 is None or not pos_has_prefix) else prefix + pos2[:n_prefix]
    return x


def resample_pos_embed(pos, new_size, old_size, num_prefix_tokens):
    """resample_abs_pos_embed (timm/layers/pos_embed.py:19-59): the grid part of a learned position table [1, P + gh*gw, D]
    resized to `new_size` by an antialiased bicubic interpolation in fp32 (a torch library call on a table of a few
    hundred rows, as in the reference); the prefix rows are kept.  Differentiable: training resamples every step."""
    if new_size[0] * new_size[1] + num_prefix_tokens == pos.shape[1] and new_size[0] == new_size[1]:
        return pos
    prefix, grid = pos[:, :num_prefix_tokens], pos[:, num_prefix_tokens:]
    D = pos.shape[-1]
    g = grid.float().reshape(1, old_size[0], old_size[1], D).permute(0, 3, 1, 2)
    g = F.interpolate(g, size=tuple(new_size), mode="bicubic", antialias=True)
    g = g.permute(0, 2, 3, 1).reshape(1, -1, D).to(pos.dtype)
    return torch.cat([prefix, g], 1) if num_prefix_tokens else g


def _dynamic_pos_embed(model, img):
    """Position table for this input size of a dynamic_img_size model (_pos_embed, vision_transformer.py:1049-1058)."""
    pe, pos = model.patch_embed, model.pos_embed
    new = (img.shape[-2] // pe.patch_size[0], img.shape[-1] // pe.patch_size[1])
    npt = 0 if getattr(model, "no_embed_class", False) else model.num_prefix_tokens
    if torch.is_grad_enabled() and pos.requires_grad:
        return resample_pos_embed(pos, new, pe.grid_size, npt)
    key = (*ops.wkey(pos), new)
    cache = getattr(model, "_b200_pos_cache", None)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            cache = (key, resample_pos_embed(pos.detach(), new, pe.grid_size, npt).contiguous())
        model._b200_pos_cache = cache
    return cache[1]


def vit_embed(model, img):
    """patch_embed + _pos_embed (+ norm_pre) of VisionTransformer.forward_features (:1238-1241)."""
    if not _is_identity(getattr(model, "patch_drop", None)):
        raise B200Error("patch dropout is not supported")
    _check_drop(getattr(model, "pos_drop", None), "pos_drop")
    pre = [t for t in (getattr(model, "cls_token", None), getattr(model, "reg_token", None)) if t is not None]
    n_prefix = sum(t.shape[1] for t in pre)
    pos = model.pos_embed
    if pos is not None and getattr(model, "dynamic_img_size", False):
        pos = _dynamic_pos_embed(model, img)
    grad = torch.is_grad_enabled() and (_needs_grad(None, model.patch_embed) or any(
        t is not None and t.requires_grad for t in pre + [pos]))
    if not grad and isinstance(img, torch.Tensor) and img.dim() == 4:
        prefix0 = None if not pre else (pre[0] if len(pre) == 1 else torch.cat(pre, 1)).reshape(n_prefix, -1)
        x = _fused_patch_tokens(model, img, pos, prefix0, n_prefix, not getattr(model, "no_embed_class", False))
        if x is not None:
            if not _is_identity(getattr(model, "norm_pre", None)):
                x = layernorm(model.norm_pre, x)
            return x
    proj = patch_embed_forward(model.patch_embed, img)
    B, G, D = proj.shape
    prefix = None
    if pre:
        prefix = (pre[0] if len(pre) == 1 else torch.cat(pre, 1)).reshape(n_prefix, D).contiguous()
    if pos is None and n_prefix == 0:
        x = proj
    elif torch.is_grad_enabled() and (proj.requires_grad or any(t is not None and t.requires_grad for t in pre + [pos])):
        from .autograd import TokensFn
        prefix_t = None if not pre else (pre[0] if len(pre) == 1 else torch.cat(pre, 1))
        x = TokensFn.apply(proj, prefix_t, pos, not getattr(model, "no_embed_class", False))
    else:
        x = ops.assemble_tokens(proj.view(B * G, D), prefix, None if pos is None else pos.contiguous(), B, G, n_prefix, D,
                                pos_has_prefix=not getattr(model, "no_embed_class", False))
    if not _is_identity(getattr(model, "norm_pre", None)):
        x = layernorm(model.norm_pre, x)
    return x


def vit_forward_features(model, img, kv_len=None, final_norm=True):
    """VisionTransformer.forward_features (timm/models/vision_transformer.py:1231-1253)."""
    x = vit_embed(model, img)
    if getattr(model, "grad_checkpointing", False) and torch.is_grad_enabled() and x.requires_grad:
        # checkpoint_seq(self.blocks, x) (vision_transformer.py:1247-1248, _manipulate.py:213-286, every=1): activations
        # of a block are dropped after its forward and the block's kernels run again in the backward; the non-reentrant
        # checkpoint restores the CUDA RNG state first, so the DropPath draw (functional._drop_path_factor) repeats
        from torch.utils.checkpoint import checkpoint
        for blk in model.blocks:
            x = checkpoint(block_forward, blk, x, kv_len, use_reentrant=False)
    else:
        carry = LnCarry()
        for blk in model.blocks:
            x = block_forward(blk, x, kv_len, carry)
    if final_norm and not _is_identity(model.norm):
        x = layernorm(model.norm, x)
    return x


def _linear(lin, x):
    """nn.Linear on [..., K] bf16 activations through LinearFn (training) or the GEMM (inference)."""
    if _needs_grad(x, lin):
        from .autograd import LinearFn
        return LinearFn.apply(x, _w(lin), lin.bias)
    shape = x.shape
    x2 = _bf16_2d(x, shape[-1])
    if lin.weight.shape[0] % 8:
        _w(lin)
        return _padded_out_linear(lin, x2).reshape(*shape[:-1], -1)
    return ops.linear_fwd(x2, _w(lin), lin.bias).view(*shape[:-1], -1)


def sdpa_packed(qkv, num_heads, scale, kv_len=None):
    """Attention on a packed bf16 [B, N, 3 * H * 64] buffer (q | k | v thirds, heads inside) -> [B, N, H * 64]."""
    B, N, C3 = qkv.shape
    if torch.is_grad_enabled() and qkv.requires_grad:
        from .autograd import SdpaFn
        return SdpaFn.apply(qkv, kv_len, num_heads, scale)
    return ops.attention_fwd(qkv.contiguous(), B, N, num_heads, C3 // 3 // num_heads, scale, kv_len)[0]


def attention_pool_forward(pool, x):
    """AttentionPoolLatent.forward (timm/layers/attention_pool.py:95-128; SigLIP's 'map' head with one latent, or several):
    latents -> q projection; tokens (+ optional position embedding) -> kv projection; L-query attention on the
    self-attention kernels (autograd.AttnPoolFn); proj; x + mlp(norm(x)); 'token' / 'avg' / no pooling over the latents."""
    if pool.pool not in ("token", "avg", ""):
        raise B200Error(f"AttentionPoolLatent: pool={pool.pool!r} is not covered")
    qk_norm = not _is_identity(pool.q_norm) or not _is_identity(pool.k_norm)
    H, hd = pool.num_heads, pool.head_dim
    P = head_pad(hd)
    if hd != 64 and qk_norm:
        raise B200Error("qk_norm with head_dim != 64 is not covered")
    _check_drop(getattr(pool, "proj_drop", None), "proj_drop")
    B, N, D = x.shape
    Dp = H * P
    L = pool.latent_len
    if L > N:
        raise B200Error("AttentionPoolLatent: more latent queries than tokens is not covered")
    _bf16_2d(x, D)
    if getattr(pool, "pos_embed", None) is not None:
        x = x + pool.pos_embed.unsqueeze(0).to(x.dtype)                  # attention_pool.py:97-99, a [N, D] table
    latent = pool.latent.reshape(L, D)
    latent = latent if latent.dtype == torch.bfloat16 else latent.to(torch.bfloat16)
    wq, bq = head_padded_linear_in(pool.q, 1, H, hd)
    wkv, bkv = head_padded_linear_in(pool.kv, 2, H, hd)
    if _needs_grad(x, pool):
        if qk_norm:
            raise B200Error("AttentionPoolLatent with qk_norm: the backward is not covered")
        from .autograd import AttnPoolFn, LinearFn
        q = LinearFn.apply(latent, wq, bq)
        o = AttnPoolFn.apply(x, q, wkv, bkv, H, pool.scale)
    else:
        q = ops.linear_fwd(latent.contiguous(), wq, bq)
        qkv = torch.zeros(B * N, 3 * Dp, dtype=torch.bfloat16, device=x.device)
        ops.linear_fwd(_bf16_2d(x, D), wkv, bkv, out=qkv[:, Dp:])
        qkv.view(B, N, 3 * Dp)[:, :L, :Dp] = q.reshape(1, L, Dp)
        if qk_norm:       # q_norm / k_norm over head_dim (:107): in place on the q and k thirds (the unused query rows too)
            (qw, qb, qe), (kw, kb, ke) = _ln_params(pool.q_norm), _ln_params(pool.k_norm)
            if float(qe) != float(ke) or (qb is None) != (kb is None):
                raise B200Error("q_norm / k_norm need equal eps and affine parameters")
            ops.qk_norm_fwd(qkv, B * N, H, qw, qb, kw, kb, qe)
        o = ops.attention_fwd(qkv, B, N, H, P, pool.scale, None)[0][:, :L].contiguous()
    if not _is_identity(pool.proj):
        wp = head_padded_linear_out(pool.proj, H, hd)
        if _needs_grad(o, pool.proj):
            from .autograd import LinearFn
            o = LinearFn.apply(o, wp, pool.proj.bias)
        else:
            o = ops.linear_fwd(o.reshape(-1, Dp), wp, pool.proj.bias).view(B, L, -1)
    elif hd != P:
        o = o.view(B, L, H, P)[..., :hd].reshape(B, L, D)
    if pool.mlp is not None:
        h = o if _is_identity(pool.norm) else layernorm(pool.norm, o)
        o = o + mlp_forward(pool.mlp, h)
    if pool.pool == "token":
        return o[:, 0]
    if pool.pool == "avg":
        return o.float().mean(1).to(o.dtype)
    return o


def attention_pool_prr_forward(pool, x):
    """AttentionPoolPrr.forward (timm/layers/attention_pool.py:172-195): pre_norm -> parameter-free multi-head self-attention
    (q = k = v = the normalised tokens) -> post_norm -> 'token' / 'avg' pooling.  The tokens are packed three times into
    the [B, N, 3, H, 64] layout the attention kernels read (a data-movement cat; its autograd adds dq + dk + dv)."""
    B, N, D = x.shape
    H, hd = pool.num_heads, pool.head_dim
    P = head_pad(hd)
    _bf16_2d(x, D)
    h = x if _is_identity(pool.pre_norm) else layernorm(pool.pre_norm, x)
    hp = h if hd == P else F.pad(h.reshape(B, N, H, hd), (0, P - hd)).reshape(B, N, H * P)
    o = sdpa_packed(torch.cat([hp, hp, hp], dim=-1), H, pool.scale)
    if hd != P:
        o = o.view(B, N, H, P)[..., :hd].reshape(B, N, D)
    o = o if _is_identity(pool.post_norm) else layernorm(pool.post_norm, o)
    if pool.pool == "token":
        return o[:, 0]
    if pool.pool == "avg":
        return o.float().mean(1).to(o.dtype)
    return o


def vit_pool(model, x):
    """VisionTransformer.pool / global_pool_nlc (:686-708, :1254-1277) for token / avg pooling."""
    if getattr(model, "attn_pool", None) is not None:                     # :1265-1269
        kind = type(model.attn_pool).__name__
        if kind not in ("AttentionPoolLatent", "AttentionPoolPrr"):
            raise B200Error(f"{kind} pooling is not supported")
        if not getattr(model, "pool_include_prefix", False) and model.num_prefix_tokens:
            x = x[:, model.num_prefix_tokens:].contiguous()
        return attention_pool_forward(model.attn_pool, x) if kind == "AttentionPoolLatent" else \
            attention_pool_prr_forward(model.attn_pool, x)
    pool = model.global_pool
    if pool == "token":
        return x[:, 0]
    if pool == "avg":
        xs = x if getattr(model, "pool_include_prefix", False) else x[:, model.num_prefix_tokens:]
        return xs.float().mean(1).to(x.dtype)
    if not pool:
        return x
    raise B200Error(f"global_pool={pool!r} is not supported")


def vit_forward_head(model, x, pre_logits=False):
    """VisionTransformer.forward_head (:1279-1292)."""
    x = vit_pool(model, x)
    if not _is_identity(model.fc_norm):
        x = layernorm(model.fc_norm, x.contiguous())
    _check_drop(model.head_drop, "head_drop")
    if pre_logits or _is_identity(model.head):
        return x
    if _needs_grad(x, model.head):
        from .autograd import LinearFn
        _bf16_2d(x, x.shape[-1])
        return LinearFn.apply(x, _w(model.head), model.head.bias)
    x2 = _bf16_2d(x, x.shape[-1])
    if model.head.weight.shape[0] % 8:
        _w(model.head)
        return _padded_out_linear(model.head, x2).reshape(*x.shape[:-1], -1)
    return ops.linear_fwd(x2, _w(model.head), model.head.bias).view(*x.shape[:-1], -1)


def vit_forward(model, img, kv_len=None):
    """VisionTransformer.forward (:1294-1302).  With token pooling the final LayerNorm is applied to the
    class-token rows only (LN is per token, so LN(x)[:, 0] == LN(x[:, 0]))."""
    token_pool = model.global_pool == "token" and getattr(model, "attn_pool", None) is None
    x = vit_forward_features(model, img, kv_len, final_norm=not token_pool)
    if token_pool and not _is_identity(model.norm):
        x = layernorm(model.norm, x[:, :1].contiguous())
    return vit_forward_head(model, x)
