"""Training path: torch.autograd.Function wrappers whose forward AND backward run on libb200timm kernels.

The reference has no backward code for this path -- gradients come from PyTorch autograd formulas of
F.linear / F.layer_norm / SDPA / GELU (SURVEY.md 8a "Backward").  These Functions own the backward definitions;
they are validated against autograd of the oracle (tests/test_gpu_train.py) and against the reference's own
backward golden (tests/golden/vit_g1_default.pt "grads").

Per transformer block (vision_transformer.py:206-214) the saved tensors are: x, LN1 stats + output, qkv, attention
output + log-sum-exp, x after the attention branch, LN2 stats + output, fc1 pre-activation and its GELU (16 N*D
units).  Backward = 4 dgrad GEMMs (GELU' fused into fc2's dgrad epilogue), 4 wgrad GEMMs (fp32 split-K through TMA
reduce-add), attention backward, 2 LayerNorm backwards with the residual gradient folded in, 4 bias column sums.
Parameters stay ordinary nn.Parameter leaves, so DDP's reducer hooks, optimizers and EMA work unchanged.
"""
import torch

from . import ops
from ._lib import B200Error


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def _to_param_dtype(g32, like):
    """fp32 gradient -> a fresh tensor with the parameter's dtype and shape (kernel conversion, no torch math)."""
    if like.dtype == torch.float32:
        return g32.view(like.shape)
    out = torch.empty(like.shape, dtype=like.dtype, device=g32.device)
    n = g32.numel()
    if n % 4 == 0:
        ops.convert_f32(g32.reshape(-1), out.view(-1))
    else:  # tiny odd-sized parameter (e.g. a 10-class head bias): pad to a multiple of 4
        pad = torch.zeros((n + 3) // 4 * 4, dtype=torch.float32, device=g32.device)
        pad[:n] = g32.reshape(-1)
        tmp = torch.empty(pad.numel(), dtype=like.dtype, device=g32.device)
        ops.convert_f32(pad, tmp)
        out.view(-1).copy_(tmp[:n])
    return out


def _bias_grad(dy2d, bias):
    if bias is None:
        return None
    return _to_param_dtype(ops.colsum(dy2d), bias)


def _wgrad(dy2d, x2d, weight):
    return _to_param_dtype(ops.linear_wgrad(dy2d, x2d), weight)


class LayerNormFn(torch.autograd.Function):
    """LayerNorm.forward (timm/layers/norm.py:85-90) with a kernel backward."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        y, mean, rstd = ops.layernorm_fwd(x, weight, bias, eps, save_stats=True)
        ctx.save_for_backward(x, weight, mean, rstd)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, mean, rstd = ctx.saved_tensors
        dx, dw, db = ops.layernorm_bwd(_c(dy), x, weight, mean, rstd)
        return dx, _to_param_dtype(dw, weight), (_to_param_dtype(db, weight) if ctx.has_bias else None), None


class LinearFn(torch.autograd.Function):
    """nn.Linear on [..., K] activations (classifier head, vision_transformer.py:1292).  out_features that are
    not a multiple of 8 run on zero-padded weight / bias copies (16-byte TMA rows) and are sliced back."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        shape = x.shape
        x2 = _c(x.reshape(-1, shape[-1]))
        N, K = weight.shape
        npad = (N + 7) // 8 * 8
        w, b = weight, bias
        if npad != N:
            w = torch.zeros(npad, K, dtype=weight.dtype, device=weight.device)
            w[:N] = weight
            if bias is not None:
                b = torch.zeros(npad, dtype=bias.dtype, device=bias.device)
                b[:N] = bias
        y = ops.linear_fwd(x2, _c(w), b)
        ctx.save_for_backward(x2, w, weight, bias)
        ctx.shape = shape
        return y[:, :N].reshape(*shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w, weight, bias = ctx.saved_tensors
        N = weight.shape[0]
        npad = w.shape[0]
        dy2 = dy.reshape(-1, N)
        if npad != N:
            pad = torch.zeros(dy2.shape[0], npad, dtype=dy2.dtype, device=dy2.device)
            pad[:, :N] = dy2
            dy2 = pad
        dy2 = _c(dy2)
        dx = ops.linear_dgrad(dy2, w).view(ctx.shape) if ctx.needs_input_grad[0] else None
        gw = ops.linear_wgrad(dy2, x2)[:N]
        gb = None if bias is None else ops.colsum(dy2)[:N]
        return dx, _to_param_dtype(_c(gw), weight), (None if gb is None else _to_param_dtype(_c(gb), bias))


class MlpFn(torch.autograd.Function):
    """Mlp.forward (timm/layers/mlp.py:47-54): fc2(norm(act(fc1(x)))); norm = Identity unless the Mlp was built with a
    norm_layer (`scale_mlp_norm`, mlp.py:44): then the activation backward is its own elementwise pass (b200_act_bwd)
    instead of the fc2-dgrad epilogue."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, act_kind=0, nw=None, nb=None, eps=0.0):
        shape = x.shape
        x2 = _c(x.reshape(-1, shape[-1]))
        pre = torch.empty(x2.shape[0], w1.shape[0], dtype=torch.bfloat16, device=x.device)
        act = ops.linear_fwd(x2, w1, b1, epilogue=ops.EPI_BIAS_GELU, out2=pre, act=act_kind)
        an, mean, rstd = act, None, None
        if nw is not None:
            an, mean, rstd = ops.layernorm_fwd(act, nw, nb, eps, save_stats=True)
        y = ops.linear_fwd(an, w2, b2)
        ctx.save_for_backward(x2, w1, b1, w2, b2, pre, act, an, nw, nb, mean, rstd)
        ctx.shape, ctx.act_kind = shape, act_kind
        return y.view(*shape[:-1], -1)

    @staticmethod
    def backward(ctx, dy):
        x2, w1, b1, w2, b2, pre, act, an, nw, nb, mean, rstd = ctx.saved_tensors
        dy2 = _c(dy.reshape(-1, dy.shape[-1]))
        gnw = gnb = None
        if nw is None:
            dpre = ops.linear_dgrad(dy2, w2, epilogue=ops.EPI_DGELU, aux=pre, act=ctx.act_kind)
        else:
            dan = ops.linear_dgrad(dy2, w2)
            da, dnw, dnb = ops.layernorm_bwd(dan, act, nw, mean, rstd)
            gnw = _to_param_dtype(dnw, nw)
            gnb = None if nb is None else _to_param_dtype(dnb, nb)
            dpre = ops.act_bwd(_c(da), pre, ctx.act_kind)
        dx = ops.linear_dgrad(dpre, w1).view(ctx.shape)
        return (dx, _wgrad(dpre, x2, w1), _bias_grad(dpre, b1), _wgrad(dy2, an, w2), _bias_grad(dy2, b2), None, gnw, gnb, None)


class AttentionFn(torch.autograd.Function):
    """Attention.forward (timm/layers/attention.py:108-139): qkv GEMM, SDPA, proj GEMM."""

    @staticmethod
    def forward(ctx, x, wqkv, bqkv, wproj, bproj, kv_len, num_heads, scale):
        B, N, C = x.shape
        hd = wqkv.shape[0] // 3 // num_heads
        x2 = _c(x.reshape(B * N, C))
        qkv = ops.linear_fwd(x2, wqkv, bqkv)
        o, lse = ops.attention_fwd(qkv, B, N, num_heads, hd, scale, kv_len, need_lse=True)
        y = ops.linear_fwd(o.view(B * N, -1), wproj, bproj)
        ctx.save_for_backward(x2, wqkv, bqkv, wproj, bproj, qkv, o, lse, kv_len)
        ctx.meta = (B, N, num_heads, hd, scale)
        return y.view(B, N, -1)

    @staticmethod
    def backward(ctx, dy):
        x2, wqkv, bqkv, wproj, bproj, qkv, o, lse, kv_len = ctx.saved_tensors
        B, N, H, hd, scale = ctx.meta
        dy2 = _c(dy.reshape(B * N, -1))
        o2 = o.view(B * N, -1)
        do = ops.linear_dgrad(dy2, wproj)
        dqkv = ops.attention_bwd(qkv, o, do.view(B, N, -1), lse, B, N, H, hd, scale, kv_len).view(B * N, -1)
        dx = ops.linear_dgrad(dqkv, wqkv).view(B, N, -1)
        return (dx, _wgrad(dqkv, x2, wqkv), _bias_grad(dqkv, bqkv), _wgrad(dy2, o2, wproj), _bias_grad(dy2, bproj),
                None, None, None)


class AttentionRopeFn(torch.autograd.Function):
    """AttentionRope.forward (timm/layers/attention.py:224-282).  The saved qkv holds the ROTATED q / k (what the
    attention kernels consumed); the backward un-rotates d(q), d(k) in place (transpose of the rotation) before the
    QKV projection's dgrad / wgrad."""

    @staticmethod
    def forward(ctx, x, wqkv, bqkv, wproj, bproj, emb, kv_len, num_heads, scale, npt, half):
        from .functional import rope_qkv_projection
        B, N, C = x.shape
        hd = wqkv.shape[0] // 3 // num_heads
        x2 = _c(x.reshape(B * N, C))
        qkv = rope_qkv_projection(x2, wqkv, bqkv, emb, B, N, num_heads, npt, half)
        o, lse = ops.attention_fwd(qkv, B, N, num_heads, hd, scale, kv_len, need_lse=True)
        y = ops.linear_fwd(o.view(B * N, -1), wproj, bproj)
        ctx.save_for_backward(x2, wqkv, bqkv, wproj, bproj, qkv, o, lse, kv_len, emb)
        ctx.meta = (B, N, num_heads, hd, scale, npt, half)
        return y.view(B, N, -1)

    @staticmethod
    def backward(ctx, dy):
        x2, wqkv, bqkv, wproj, bproj, qkv, o, lse, kv_len, emb = ctx.saved_tensors
        B, N, H, hd, scale, npt, half = ctx.meta
        dy2 = _c(dy.reshape(B * N, -1))
        o2 = o.view(B * N, -1)
        do = ops.linear_dgrad(dy2, wproj)
        dqkv = ops.attention_bwd(qkv, o, do.view(B, N, -1), lse, B, N, H, hd, scale, kv_len).view(B * N, -1)
        if emb is not None:
            ops.rope_qk(dqkv, emb, B, N, H, npt, half=half, inverse=True)
        dx = ops.linear_dgrad(dqkv, wqkv).view(B, N, -1)
        return (dx, _wgrad(dqkv, x2, wqkv), _bias_grad(dqkv, bqkv), _wgrad(dy2, o2, wproj), _bias_grad(dy2, bproj),
                None, None, None, None, None, None)


class BlockFn(torch.autograd.Function):
    """Block.forward (timm/models/vision_transformer.py:206-214) incl. LayerScale (layer_scale.py:26-27) and
    DropPath (drop.py:158-175): x = x + dp1(ls1(attn(norm1(x)))); x = x + dp2(ls2(mlp(norm2(x)))).

    rs1 / rs2: per-sample DropPath factors (keep-mask / keep_prob, fp32 [B]) drawn by the caller from torch's CUDA
    generator exactly as the reference draws them, or None.  g1 / g2: LayerScale gammas or None.  When a branch has
    either, its un-scaled output is kept (GEMM out2) for the gamma gradient and its incoming gradient is scaled by
    gamma[col] * rs[row] before the projection / fc2 backward."""

    @staticmethod
    def forward(ctx, x, n1w, n1b, wqkv, bqkv, wproj, bproj, n2w, n2b, w1, b1, w2, b2, g1, g2, rs1, rs2, kv_len,
                num_heads, scale, eps1, eps2, act_kind=0):
        B, N, D = x.shape
        hd = wqkv.shape[0] // 3 // num_heads
        x0 = _c(x.reshape(B * N, D))
        h1, mean1, rstd1 = ops.layernorm_fwd(x0, n1w, n1b, eps1, save_stats=True)
        qkv = ops.linear_fwd(h1, wqkv, bqkv)
        o, lse = ops.attention_fwd(qkv, B, N, num_heads, hd, scale, kv_len, need_lse=True)
        br1 = torch.empty(B * N, D, dtype=torch.bfloat16, device=x.device) if g1 is not None else None
        x1 = ops.linear_fwd(o.view(B * N, -1), wproj, bproj, epilogue=ops.EPI_BIAS_RESIDUAL, aux=x0, gamma=g1, out2=br1,
                            row_scale=rs1, rows_per_scale=N)
        h2, mean2, rstd2 = ops.layernorm_fwd(x1, n2w, n2b, eps2, save_stats=True)
        pre = torch.empty(B * N, w1.shape[0], dtype=torch.bfloat16, device=x.device)
        act = ops.linear_fwd(h2, w1, b1, epilogue=ops.EPI_BIAS_GELU, out2=pre, act=act_kind)
        br2 = torch.empty(B * N, D, dtype=torch.bfloat16, device=x.device) if g2 is not None else None
        x2 = ops.linear_fwd(act, w2, b2, epilogue=ops.EPI_BIAS_RESIDUAL, aux=x1, gamma=g2, out2=br2, row_scale=rs2,
                            rows_per_scale=N)
        ctx.save_for_backward(x0, n1w, n1b, wqkv, bqkv, wproj, bproj, n2w, n2b, w1, b1, w2, b2, g1, g2, rs1, rs2, kv_len,
                              h1, mean1, rstd1, qkv, o, lse, x1, h2, mean2, rstd2, pre, act, br1, br2)
        ctx.meta = (B, N, D, num_heads, hd, scale)
        ctx.act_kind = act_kind
        return x2.view(B, N, D)

    @staticmethod
    def backward(ctx, dy):
        (x0, n1w, n1b, wqkv, bqkv, wproj, bproj, n2w, n2b, w1, b1, w2, b2, g1, g2, rs1, rs2, kv_len, h1, mean1, rstd1, qkv,
         o, lse, x1, h2, mean2, rstd2, pre, act, br1, br2) = ctx.saved_tensors
        B, N, D, H, hd, scale = ctx.meta
        dx2 = _c(dy.reshape(B * N, D))
        # All parameter gradients of the block live in ONE zeroed fp32 workspace (wgrad split-K reduce-adds, bias
        # column sums, LayerNorm dgamma/dbeta and LayerScale dgamma write straight into views of it) and are converted
        # to the parameters' dtype with ONE kernel: 1 memset + 1 convert per block instead of ~20 of each.
        params = (n1w, n1b, wqkv, bqkv, wproj, bproj, n2w, n2b, w1, b1, w2, b2, g1, g2)
        sizes = [0 if p is None else p.numel() for p in params]
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + (n + 7) // 8 * 8)      # 16-byte aligned bf16 slices (the layout of grad_sync's block slots)
        ws = torch.zeros(offs[-1], dtype=torch.float32, device=dx2.device)
        v = [None if p is None else ws[o_:o_ + n].view(p.shape) for p, o_, n in zip(params, offs, sizes)]
        # ---- MLP branch:  x2 = x1 + rs2 * g2 * (fc2(GELU(fc1(LN2(x1)))) + b2)
        dbr2 = dx2
        if g2 is not None or rs2 is not None:
            if g2 is not None:
                ops.colsum_mul(dx2, br2, rs2, N, out=v[13], accumulate=True)          # dgamma2
            dbr2 = ops.scale_rows_cols(dx2, g2, rs2, N)                               # gradient w.r.t. the branch output
        # dY W2 * act'(pre); the bias gradient of fc1 (column sums of dpre) rides in this GEMM's epilogue
        fused_db1 = b1 is not None and ops.dgrad_colsum_fused(B * N, w2.shape[1])
        dpre = ops.linear_dgrad(dbr2, w2, epilogue=ops.EPI_DGELU, aux=pre, act=ctx.act_kind, colsum=v[9] if fused_db1 else None)
        ops.linear_wgrad(dbr2, act, out=v[10])
        if b2 is not None:
            ops.colsum(dbr2, out=v[11], accumulate=True)
        dh2 = ops.linear_dgrad(dpre, w1)
        ops.linear_wgrad(dpre, h2, out=v[8])
        if b1 is not None and not fused_db1:
            ops.colsum(dpre, out=v[9], accumulate=True)
        dx1, _, _ = ops.layernorm_bwd(dh2, x1, n2w, mean2, rstd2, dres=dx2, dw_out=v[6], db_out=v[7])
        # ---- attention branch:  x1 = x0 + rs1 * g1 * (proj(SDPA(qkv(LN1(x0)))) + bp)
        dbr1 = dx1
        if g1 is not None or rs1 is not None:
            if g1 is not None:
                ops.colsum_mul(dx1, br1, rs1, N, out=v[12], accumulate=True)          # dgamma1
            dbr1 = ops.scale_rows_cols(dx1, g1, rs1, N)
        do = ops.linear_dgrad(dbr1, wproj)
        ops.linear_wgrad(dbr1, o.view(B * N, -1), out=v[4])
        if bproj is not None:
            ops.colsum(dbr1, out=v[5], accumulate=True)
        fused_dbqkv = bqkv is not None and ops.attn_bwd_bias_fused(N, hd)       # the qkv bias gradient rides in the attention backward
        dqkv = ops.attention_bwd(qkv, o, do.view(B, N, -1), lse, B, N, H, hd, scale, kv_len,
                                 dbias=v[3] if fused_dbqkv else None).view(B * N, -1)
        dh1 = ops.linear_dgrad(dqkv, wqkv)
        ops.linear_wgrad(dqkv, h1, out=v[2])
        if bqkv is not None and not fused_dbqkv:
            ops.colsum(dqkv, out=v[3], accumulate=True)
        dx0, _, _ = ops.layernorm_bwd(dh1, x0, n1w, mean1, rstd1, dres=dx1, dw_out=v[0], db_out=v[1])
        # data parallel with b200timm.grad_sync.PeerGradSync: the workspace is scaled by 1 / world_size, rounded to bf16 and
        # added into EVERY rank's flat gradient buffer by one kernel (multimem.red over the NVLink multicast address) --
        # the gradient all-reduce happens here, block by block, under the backward GEMMs of the next block; p.grad of the
        # parameters are views of that buffer, so nothing is returned to autograd for them
        from .grad_sync import active_sink
        sink = active_sink()
        slot = None if sink is None else sink.slot_of(params)
        if slot is not None and slot[1] == offs[-1]:
            sink.push_block(ws, slot[0], slot[1])
            return (dx0.view(B, N, D), *([None] * 14), None, None, None, None, None, None, None, None, None)
        # bf16 parameters (and the bf16 casts of fp32 master weights) take their gradient from one bf16 conversion of the
        # whole workspace; fp32 parameters (AMP: biases, LayerNorm, LayerScale) take the fp32 views directly
        if all(p is None or p.dtype == torch.float32 for p in params):
            grads = v
        else:
            gbuf = torch.empty(offs[-1], dtype=torch.bfloat16, device=dx2.device)
            ops.convert_f32(ws, gbuf)
            grads = [None if p is None else (v[i] if p.dtype == torch.float32 else gbuf[o_:o_ + n].view(p.shape))
                     for i, (p, o_, n) in enumerate(zip(params, offs, sizes))]
        return (dx0.view(B, N, D), *grads, None, None, None, None, None, None, None, None, None)


class PatchEmbedFn(torch.autograd.Function):
    """PatchEmbed.forward (timm/layers/patch_embed.py:117-142): patch gather + GEMM; no image gradient."""

    @staticmethod
    def forward(ctx, img, weight, bias, patch):
        B, C, H, W = img.shape
        D = weight.shape[0]
        K = C * patch * patch
        kpad = (K + 7) // 8 * 8
        w2 = weight.reshape(D, K)
        if kpad != K:
            wp = torch.zeros(D, kpad, dtype=weight.dtype, device=weight.device)
            wp[:, :K] = w2
            w2 = wp
        patches = ops.patchify(_c(img), patch, kpad)
        y = ops.gemm(patches, _c(w2), patches.shape[0], D, kpad, bias=bias)
        ctx.save_for_backward(patches, weight, bias)
        ctx.K = K
        return y.view(B, (H // patch) * (W // patch), D)

    @staticmethod
    def backward(ctx, dy):
        patches, weight, bias = ctx.saved_tensors
        dy2 = _c(dy.reshape(-1, dy.shape[-1]))
        gw = ops.linear_wgrad(dy2, patches)            # [D, Kpad] fp32
        if gw.shape[1] != ctx.K:
            gw = gw[:, :ctx.K].contiguous()
        return None, _to_param_dtype(gw, weight), _bias_grad(dy2, bias), None


class TokensFn(torch.autograd.Function):
    """_pos_embed (vision_transformer.py:1038-1075): prefix-token concat + position-embedding add."""

    @staticmethod
    def forward(ctx, proj, prefix, pos, pos_has_prefix):
        B, G, D = proj.shape
        n_prefix = 0 if prefix is None else prefix.shape[-2]
        pre2 = None if prefix is None else _c(prefix.reshape(n_prefix, D))
        pos2 = None if pos is None else _c(pos.reshape(-1, D))
        tokens = ops.assemble_tokens(_c(proj.reshape(B * G, D)), pre2, pos2, B, G, n_prefix, D, pos_has_prefix)
        ctx.meta = (B, G, D, n_prefix, pos_has_prefix)
        ctx.shapes = (None if prefix is None else (prefix.shape, prefix.dtype), None if pos is None else (pos.shape, pos.dtype))
        return tokens

    @staticmethod
    def backward(ctx, dtok):
        B, G, D, n_prefix, pos_has_prefix = ctx.meta
        dtok = _c(dtok)
        dproj = dtok[:, n_prefix:, :].contiguous() if n_prefix else dtok
        s = ops.colsum(dtok.view(B, (G + n_prefix) * D)).view(G + n_prefix, D)     # sum over the batch, fp32
        dprefix = dpos = None
        if ctx.shapes[0] is not None:
            shape, dtype = ctx.shapes[0]
            dprefix = s[:n_prefix].to(dtype).view(shape)
        if ctx.shapes[1] is not None:
            shape, dtype = ctx.shapes[1]
            dpos = (s if pos_has_prefix else s[n_prefix:]).to(dtype).view(shape)
        return dproj, dprefix, dpos, None


class NaflexEmbedFn(torch.autograd.Function):
    """NaFlexEmbeds.forward in NaFlex mode (timm/models/naflexvit.py:857-960): linear patch projection + per-sample
    position rows + prefix tokens.  `table` is the concatenated (already interpolated) position rows: its gradient is
    returned to autograd, which carries it through the bicubic resize to `pos_embed`."""

    @staticmethod
    def forward(ctx, x2, weight, bias, prefix, table, pos_off, pos_len, B, N):
        D, K = weight.shape
        kpad = x2.shape[1]
        w2 = weight
        if kpad != K:      # 16-byte row pitch for TMA (p=14: 588 -> 592); x2 arrives zero-padded already
            w2 = torch.zeros(D, kpad, dtype=weight.dtype, device=weight.device)
            w2[:, :K] = weight
        n_prefix = 0 if prefix is None else prefix.shape[0]
        proj = ops.gemm(x2, _c(w2), B * N, D, kpad, bias=bias)
        tokens = ops.naflex_assemble(proj, None if prefix is None else _c(prefix), table, pos_off, pos_len, B, N,
                                     n_prefix, D)
        ctx.save_for_backward(x2, weight, bias, pos_off, pos_len)
        ctx.meta = (B, N, D, K, n_prefix, 0 if table is None else table.shape[0],
                    None if prefix is None else prefix.dtype, None if table is None else table.dtype)
        return tokens

    @staticmethod
    def backward(ctx, dtok):
        x2, weight, bias, pos_off, pos_len = ctx.saved_tensors
        B, N, D, K, n_prefix, rows, pre_dt, tab_dt = ctx.meta
        dproj, dprefix, dtable = ops.naflex_assemble_bwd(_c(dtok), pos_off, pos_len, rows, B, N, n_prefix, D)
        gw = ops.linear_wgrad(dproj, x2)               # [D, Kpad] fp32
        if gw.shape[1] != K:
            gw = gw[:, :K].contiguous()
        if dprefix is not None:
            dprefix = _to_param_dtype(dprefix, torch.empty(dprefix.shape, dtype=pre_dt, device="meta"))
        if dtable is not None:
            dtable = _to_param_dtype(dtable, torch.empty(dtable.shape, dtype=tab_dt, device="meta"))
        return None, _to_param_dtype(gw, weight), _bias_grad(dproj, bias), dprefix, dtable, None, None, None, None


class GluMlpFn(torch.autograd.Function):
    """GluMlp / SwiGLU forward (timm/layers/mlp.py:104-113, 153-162): fc2(norm(act(gate) * value)) with the gate and the
    value as the two halves of one fc1 GEMM.  `nw is None` = no inner norm."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, nw, nb, eps, gate_first, act):
        shape = x.shape
        x2 = _c(x.reshape(-1, shape[-1]))
        h = ops.linear_fwd(x2, w1, b1)
        a = ops.glu_fwd(h, gate_first, act)
        mean = rstd = None
        an = a
        if nw is not None:
            an, mean, rstd = ops.layernorm_fwd(a, nw, nb, eps, save_stats=True)
        y = ops.linear_fwd(an, w2, b2)
        ctx.save_for_backward(x2, w1, b1, w2, b2, h, a, an, nw, nb, mean, rstd)
        ctx.meta = (shape, gate_first, act)
        return y.view(*shape[:-1], -1)

    @staticmethod
    def backward(ctx, dy):
        x2, w1, b1, w2, b2, h, a, an, nw, nb, mean, rstd = ctx.saved_tensors
        shape, gate_first, act = ctx.meta
        dy2 = _c(dy.reshape(-1, dy.shape[-1]))
        da = ops.linear_dgrad(dy2, w2)
        gnw = gnb = None
        if nw is not None:
            da, dnw, dnb = ops.layernorm_bwd(da, a, nw, mean, rstd)
            gnw = _to_param_dtype(dnw, nw)
            gnb = None if nb is None else _to_param_dtype(dnb, nb)
        dh = ops.glu_bwd(h, _c(da), gate_first, act)
        dx = ops.linear_dgrad(dh, w1).view(shape)
        return (dx, _wgrad(dh, x2, w1), _bias_grad(dh, b1), _wgrad(dy2, an, w2), _bias_grad(dy2, b2), gnw, gnb, None, None,
                None)


class AttentionNormFn(torch.autograd.Function):
    """Attention.forward with qk_norm and / or scale_norm (timm/layers/attention.py:108-139): qkv GEMM, per-head
    normalisation of q and k (in place on the qkv buffer), SDPA, normalisation of the merged heads, proj GEMM.
    qk = (q_weight, q_bias, k_weight, k_bias, eps) or None; sn = (weight, bias, eps) or None (eps may be an RmsEps)."""

    @staticmethod
    def forward(ctx, x, wqkv, bqkv, wproj, bproj, qw, qb, kw, kb, nw, nb, kv_len, num_heads, scale, qk_eps, sn_eps,
                emb=None, npt=0, half=False):
        """emb (optional): rotary table, fp32 [N - npt, 128] or per-sample [B, N - npt, 128]; q / k are rotated AFTER their
        normalisation (AttentionRope / EvaAttention order, attention.py:241-253); the saved qkv holds the rotated q / k."""
        B, N, C = x.shape
        hd = wqkv.shape[0] // 3 // num_heads
        x2 = _c(x.reshape(B * N, C))
        qkv = ops.linear_fwd(x2, wqkv, bqkv)
        raw = qmean = qrstd = None
        if qw is not None:
            raw, qmean, qrstd = ops.qk_norm_fwd(qkv, B * N, num_heads, qw, qb, kw, kb, qk_eps, save=True)
        if emb is not None:
            ops.rope_qk(qkv, emb, B, N, num_heads, npt, half=half)
        o, lse = ops.attention_fwd(qkv, B, N, num_heads, hd, scale, kv_len, need_lse=True)
        o2 = o.view(B * N, -1)
        on, nmean, nrstd = o2, None, None
        if nw is not None:
            on, nmean, nrstd = ops.layernorm_fwd(o2, nw, nb, sn_eps, save_stats=True)
        y = ops.linear_fwd(on, wproj, bproj)
        ctx.save_for_backward(x2, wqkv, bqkv, wproj, bproj, qkv, o, lse, kv_len, raw, qmean, qrstd, qw, qb, kw, kb, on, nmean,
                              nrstd, nw, nb, emb)
        ctx.meta = (B, N, num_heads, hd, scale, npt, half)
        return y.view(B, N, -1)

    @staticmethod
    def backward(ctx, dy):
        (x2, wqkv, bqkv, wproj, bproj, qkv, o, lse, kv_len, raw, qmean, qrstd, qw, qb, kw, kb, on, nmean, nrstd, nw,
         nb, emb) = ctx.saved_tensors
        B, N, H, hd, scale, npt, half = ctx.meta
        dy2 = _c(dy.reshape(B * N, -1))
        o2 = o.view(B * N, -1)
        do = ops.linear_dgrad(dy2, wproj)
        gnw = gnb = None
        if nw is not None:
            do, dnw, dnb = ops.layernorm_bwd(do, o2, nw, nmean, nrstd)
            gnw = _to_param_dtype(dnw, nw)
            gnb = None if nb is None else _to_param_dtype(dnb, nb)
        dqkv = ops.attention_bwd(qkv, o, _c(do).view(B, N, -1), lse, B, N, H, hd, scale, kv_len).view(B * N, -1)
        if emb is not None:
            ops.rope_qk(dqkv, emb, B, N, H, npt, half=half, inverse=True)
        gq = [None] * 4
        if qw is not None:
            dp = ops.qk_norm_bwd(dqkv, raw, qmean, qrstd, qw, kw, B * N, H)
            gq = [_to_param_dtype(dp[0].contiguous(), qw), None if qb is None else _to_param_dtype(dp[1].contiguous(), qb),
                  _to_param_dtype(dp[2].contiguous(), kw), None if kb is None else _to_param_dtype(dp[3].contiguous(), kb)]
        dx = ops.linear_dgrad(dqkv, wqkv).view(B, N, -1)
        return (dx, _wgrad(dqkv, x2, wqkv), _bias_grad(dqkv, bqkv), _wgrad(dy2, on, wproj), _bias_grad(dy2, bproj),
                gq[0], gq[1], gq[2], gq[3], gnw, gnb, None, None, None, None, None, None, None, None)


class AttnPoolFn(torch.autograd.Function):
    """The attention of AttentionPoolLatent (timm/layers/attention_pool.py:95-117): L learned latent queries per image attend
    over the N tokens.  It runs on the self-attention kernels: the kv projection is written straight into the k / v columns
    of a [B, N, 3, H, 64] buffer, the projected latents become query rows 0 .. L-1 of every image (the other query rows are
    zero and their outputs are never read), and output rows 0 .. L-1 are the pooled vectors.  The backward feeds d(out) as
    rows 0 .. L-1 of an otherwise zero gradient: rows with zero upstream gradient contribute nothing to dK / dV."""

    @staticmethod
    def forward(ctx, x, q, wkv, bkv, num_heads, scale):
        B, N, D = x.shape
        Dp = wkv.shape[0] // 2                        # H * 64: the (head-padded) width of q, k and v
        L = q.numel() // Dp
        if L > N:
            raise B200Error("AttentionPoolLatent: more latent queries than tokens is not covered")
        x2 = _c(x.reshape(B * N, D))
        qkv = torch.zeros(B * N, 3 * Dp, dtype=torch.bfloat16, device=x.device)
        ops.linear_fwd(x2, wkv, bkv, out=qkv[:, Dp:])
        qkv.view(B, N, 3 * Dp)[:, :L, :Dp] = q.reshape(1, L, Dp)
        o, lse = ops.attention_fwd(qkv, B, N, num_heads, Dp // num_heads, scale, None, need_lse=True)
        ctx.save_for_backward(x2, wkv, bkv, qkv, o, lse)
        ctx.meta = (B, N, D, Dp, L, num_heads, scale, q.shape, q.dtype)
        return o[:, :L].contiguous()

    @staticmethod
    def backward(ctx, dy):
        x2, wkv, bkv, qkv, o, lse = ctx.saved_tensors
        B, N, D, Dp, L, H, scale, qshape, qdtype = ctx.meta
        dout = torch.zeros(B, N, Dp, dtype=torch.bfloat16, device=dy.device)
        dout[:, :L] = dy
        dqkv = ops.attention_bwd(qkv, o, dout, lse, B, N, H, Dp // H, scale, None).view(B * N, 3 * Dp)
        dq = dqkv.view(B, N, 3 * Dp)[:, :L, :Dp].float().sum(0).to(qdtype).reshape(qshape)
        dkv = dqkv[:, Dp:].contiguous()                     # once per step, on the pooling head only
        dx = ops.linear_dgrad(dkv, wkv).view(B, N, D)
        return dx, dq, _wgrad(dkv, x2, wkv), _bias_grad(dkv, bkv), None, None


class SdpaFn(torch.autograd.Function):
    """softmax(q k^T * scale) v on a packed [B, N, 3 * H * 64] q | k | v buffer -> [B, N, H * 64]; forward and backward on the
    attention kernels.  For callers whose projections are not the fused nn.Linear of Attention (the parameter-free
    self-attention of AttentionPoolPrr, attention_pool.py:172-187; EvaAttention with separate q / k / v projections)."""

    @staticmethod
    def forward(ctx, qkv, kv_len, num_heads, scale):
        B, N, C3 = qkv.shape
        hd = C3 // 3 // num_heads
        qkv = _c(qkv)
        o, lse = ops.attention_fwd(qkv, B, N, num_heads, hd, scale, kv_len, need_lse=True)
        ctx.save_for_backward(qkv, o, lse, kv_len)
        ctx.meta = (B, N, num_heads, hd, scale)
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, o, lse, kv_len = ctx.saved_tensors
        B, N, H, hd, scale = ctx.meta
        return ops.attention_bwd(qkv, o, _c(do), lse, B, N, H, hd, scale, kv_len), None, None, None
