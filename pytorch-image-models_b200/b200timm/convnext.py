"""Mirror of timm's ConvNeXt (timm/models/convnext.py: ConvNeXtBlock :117-213, ConvNeXtStage :216-307,
ConvNeXt :338-629, NormMlpClassifierHead timm/layers/classifier.py:145-220) for the hot-path configuration:
patch stem, LayerNorm, nn.Linear MLP with exact GELU, LayerScale `gamma`, avg-pool head.  Same module / parameter
names as timm (``stem.0/1``, ``stages.i.downsample.0/1``, ``stages.i.blocks.j.{gamma,conv_dw,norm,mlp.fc1,mlp.fc2}``,
``head.norm``, ``head.fc``), so timm checkpoints load with strict=True.

B200 execution (inference): the network runs channels-last end to end -- activations are [B*H*W, C] bf16 rows --
    stem        patch gather (NCHW image) + GEMM + LayerNorm kernel
    block       depthwise 7x7 kernel (NHWC) -> LayerNorm kernel -> fc1 GEMM with fused bias+GELU ->
                fc2 GEMM with fused bias, LayerScale gamma and the shortcut add           (convnext.py:197-213)
    downsample  LayerNorm kernel -> 2x2 gather -> GEMM (conv weight permuted once to [Cout, ky, kx, Cin])
    head        avg-pool kernel -> LayerNorm kernel -> GEMM
so the permutes of convnext.py:205,208 and LayerNorm2d's permute/permute (norm.py:127-134) never happen.
"""
from typing import Optional, Tuple

import torch
from torch import nn

from . import ops
from ._lib import B200Error
from .layers import DropPath, Mlp


from . import functional as BF  # noqa: E402
from .functional import _w, mlp_act  # noqa: E402  (bf16 GEMM operand of a Linear / Conv2d: bf16 params as is, fp32 params cast)


def _needs_grad(module):
    return torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters())


class _BlockFn(torch.autograd.Function):
    """ConvNeXtBlock forward + backward on channels-last rows (timm/models/convnext.py:197-213).  Backward: LayerScale
    gamma gradient, fc2 / fc1 dgrad (GELU' fused) + wgrad, LayerNorm backward, then the depthwise conv backward --
    dx = dwconv(dy, spatially flipped weight) with the shortcut gradient folded in, dweight through the wgrad kernel."""

    @staticmethod
    def forward(ctx, x, wdw, bdw, nw, nb, w1, b1, w2, b2, gamma, rs, B, H, W, eps, gw=None, gb=None, geps=1e-6, act_kind=0):
        from .autograd import _c
        C = x.shape[1]
        k = wdw.shape[-1]
        wt = ops.dwconv_weight_t(wdw)
        y = ops.dwconv_nhwc(x.view(B, H, W, C), wt, bdw).view(B * H * W, C)
        h, mean, rstd = ops.layernorm_fwd(y, nw, nb, eps, save_stats=True)
        pre = torch.empty(x.shape[0], w1.shape[0], dtype=torch.bfloat16, device=x.device)
        act = ops.linear_fwd(h, w1, b1, epilogue=ops.EPI_BIAS_GELU, out2=pre, act=act_kind)
        gsq = gscale = None
        if gw is not None:          # ConvNeXt-V2: GlobalResponseNorm between GELU and fc2 (mlp.py:289-292), in place on act
            gsq, gscale = ops.grn_fwd(act, gw, gb, B, H * W, geps)
        br = torch.empty_like(x) if gamma is not None else None
        out = ops.linear_fwd(act, w2, b2, epilogue=ops.EPI_BIAS_RESIDUAL, aux=_c(x), gamma=gamma, out2=br, row_scale=rs,
                             rows_per_scale=H * W)
        ctx.save_for_backward(x, wdw, bdw, nw, nb, w1, b1, w2, b2, gamma, wt, y, mean, rstd, h, pre, act, br, rs, gw, gb, gsq,
                              gscale)
        ctx.meta = (B, H, W, C, k, geps)
        ctx.act_kind = act_kind
        return out

    @staticmethod
    def backward(ctx, dout):
        from .autograd import _c, _to_param_dtype
        (x, wdw, bdw, nw, nb, w1, b1, w2, b2, gamma, wt, y, mean, rstd, h, pre, act, br, rs, gw, gb, gsq,
         gscale) = ctx.saved_tensors
        B, H, W, C, k, geps = ctx.meta
        dout = _c(dout)
        params = (bdw, nw, nb, w1, b1, w2, b2, gamma, gw, gb)
        sizes = [0 if p is None else p.numel() for p in params]
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + (n + 3) // 4 * 4)
        kk_off = offs[-1]
        ws = torch.zeros(kk_off + k * k * C, dtype=torch.float32, device=dout.device)
        v = [None if p is None else ws[o:o + n].view(p.shape) for p, o, n in zip(params, offs, sizes)]
        dwt = ws[kk_off:].view(k, k, C)
        dbr = dout
        if gamma is not None or rs is not None:        # out = x + rs[sample] * gamma * branch  (DropPath, LayerScale)
            if gamma is not None:
                ops.colsum_mul(dout, br, rs, H * W, out=v[7], accumulate=True)
            dbr = ops.scale_rows_cols(dout, gamma, rs, H * W)
        if gw is None:
            dpre = ops.linear_dgrad(dbr, w2, epilogue=ops.EPI_DGELU, aux=pre, act=ctx.act_kind)
        else:                       # GRN + GELU backward in one pass over the hidden activations
            dact = ops.linear_dgrad(dbr, w2)
            if gb is not None:
                ops.colsum(dact, out=v[9], accumulate=True)
            dpre, _ = ops.grn_bwd(dact, pre, gsq, gscale, gw, B, H * W, geps, dweight=v[8])
        ops.linear_wgrad(dbr, act, out=v[5])
        if b2 is not None:
            ops.colsum(dbr, out=v[6], accumulate=True)
        dh = ops.linear_dgrad(dpre, w1)
        ops.linear_wgrad(dpre, h, out=v[3])
        if b1 is not None:
            ops.colsum(dpre, out=v[4], accumulate=True)
        dyc, _, _ = ops.layernorm_bwd(dh, y, nw, mean, rstd, dw_out=v[1], db_out=v[2])
        dx = ops.dwconv_nhwc(dyc.view(B, H, W, C), wt.flip(0, 1).contiguous(), None, residual=dout.view(B, H, W, C))
        ops.dwconv_nhwc_wgrad(x.view(B, H, W, C), dyc.view(B, H, W, C), k, out=dwt)
        if bdw is not None:
            ops.colsum(dyc, out=v[0], accumulate=True)
        # bf16 parameters / bf16 casts of fp32 master weights: one bf16 conversion of the workspace; fp32 parameters
        # (AMP: biases, LayerNorm, gamma): the fp32 views themselves
        if all(p is None or p.dtype == torch.float32 for p in params):
            grads = v
        else:
            gbuf = torch.empty(kk_off, dtype=torch.bfloat16, device=dout.device)
            ops.convert_f32(ws[:kk_off], gbuf)
            grads = [None if p is None else (v[i] if p.dtype == torch.float32 else gbuf[o:o + n].view(p.shape))
                     for i, (p, o, n) in enumerate(zip(params, offs, sizes))]
        g_wdw = dwt.reshape(k * k, C).t().reshape(C, 1, k, k).to(wdw.dtype)     # [k,k,C] -> Conv2d layout (49*C values)
        return (dx.view(B * H * W, C), g_wdw, grads[0], grads[1], grads[2], grads[3], grads[4], grads[5], grads[6], grads[7],
                None, None, None, None, None, grads[8], grads[9], None, None)


class _DownsampleFn(torch.autograd.Function):
    """LayerNorm2d + 2x2 / stride-2 conv of ConvNeXtStage (convnext.py:263-275, 300-302) on channels-last rows."""

    @staticmethod
    def forward(ctx, x, nw, nb, wconv, bconv, B, H, W, eps):
        C = x.shape[1]
        p = wconv.shape[-1]
        h, mean, rstd = ops.layernorm_fwd(x, nw, nb, eps, save_stats=True)
        w2 = wconv.permute(0, 2, 3, 1).reshape(wconv.shape[0], -1).contiguous()
        patches = ops.patchify_nhwc(h.view(B, H, W, C), p) if p > 1 else h
        out = ops.linear_fwd(patches, w2, bconv)
        ctx.save_for_backward(x, nw, nb, wconv, bconv, mean, rstd, patches, w2)
        ctx.meta = (B, H, W, C, p)
        return out

    @staticmethod
    def backward(ctx, dout):
        from .autograd import _c, _to_param_dtype
        x, nw, nb, wconv, bconv, mean, rstd, patches, w2 = ctx.saved_tensors
        B, H, W, C, p = ctx.meta
        dout = _c(dout)
        dpatches = ops.linear_dgrad(dout, w2)
        gw2 = ops.linear_wgrad(dout, patches)                                     # [Cout, p*p*Cin] fp32
        g_w = gw2.view(wconv.shape[0], p, p, C).permute(0, 3, 1, 2).contiguous().to(wconv.dtype)
        g_b = None if bconv is None else _to_param_dtype(ops.colsum(dout), bconv)
        dh = ops.unpatchify_nhwc(dpatches, B, H, W, C, p).view(B * H * W, C) if p > 1 else dpatches
        dx, dnw, dnb = ops.layernorm_bwd(dh, x, nw, mean, rstd)
        return dx, _to_param_dtype(dnw, nw), _to_param_dtype(dnb, nb), g_w, g_b, None, None, None, None


class _AvgPoolFn(torch.autograd.Function):
    """Global average pool over H*W (classifier.py:211-213); backward broadcasts dy / HW."""

    @staticmethod
    def forward(ctx, x, B, HW):
        ctx.meta = (B, HW, x.shape[-1])
        return ops.avgpool_nhwc(x.view(B, HW, -1))

    @staticmethod
    def backward(ctx, dy):
        B, HW, C = ctx.meta
        g = (dy.float() / HW).to(dy.dtype)
        return g[:, None, :].expand(B, HW, C).reshape(B * HW, C), None, None


class LayerNorm2d(nn.LayerNorm):
    """timm.layers.LayerNorm2d (norm.py:113-134): LayerNorm over C of an NCHW tensor.  Parameter holder here --
    the fused forward applies it on channels-last rows."""

    def __init__(self, num_channels: int, eps: float = 1e-6, affine: bool = True, device=None, dtype=None):
        super().__init__(num_channels, eps=eps, elementwise_affine=affine, device=device, dtype=dtype)

    def forward(self, x):
        raise B200Error("LayerNorm2d is applied channels-last inside ConvNeXt.forward; standalone NCHW use is not covered")


class GlobalResponseNorm(nn.Module):
    """Parameter holder of timm.layers.GlobalResponseNorm (grn.py:17-49), channels-last; applied by b200_grn_fwd."""

    def __init__(self, dim: int, eps: float = 1e-6, channels_last: bool = True, device=None, dtype=None):
        super().__init__()
        if not channels_last:
            raise B200Error("GlobalResponseNorm: channels-first is not covered")
        self.eps = eps
        self.weight = nn.Parameter(torch.zeros(dim, device=device, dtype=dtype))
        self.bias = nn.Parameter(torch.zeros(dim, device=device, dtype=dtype))


class GlobalResponseNormMlp(nn.Module):
    """timm.layers.GlobalResponseNormMlp (mlp.py:251-295), Linear / channels-last form: fc1, GELU, GRN, fc2.  It runs
    inside ConvNeXtBlock's fused forward (block_forward_nhwc), which needs the image's spatial extent."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, bias=True, drop=0.,
                 use_conv=False, device=None, dtype=None):
        super().__init__()
        if use_conv:
            raise B200Error("GlobalResponseNormMlp(use_conv=True) is not covered")
        dd = {"device": device, "dtype": dtype}
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias, **dd)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.grn = GlobalResponseNorm(hidden_features, **dd)
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias, **dd)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        raise B200Error("GlobalResponseNormMlp runs inside ConvNeXtBlock (it needs the spatial extent of each image)")


class ConvNeXtBlock(nn.Module):
    def __init__(self, in_chs: int, out_chs: Optional[int] = None, kernel_size: int = 7, stride: int = 1,
                 dilation=(1, 1), mlp_ratio: float = 4, conv_mlp: bool = False, conv_bias: bool = True,
                 use_grn: bool = False, ls_init_value: Optional[float] = 1e-6, act_layer=nn.GELU, norm_layer=None,
                 drop_path: float = 0., device=None, dtype=None):
        super().__init__()
        dd = {"device": device, "dtype": dtype}
        out_chs = out_chs or in_chs
        if conv_mlp or stride != 1 or in_chs != out_chs or tuple(dilation) != (1, 1):
            raise B200Error("ConvNeXtBlock: conv_mlp / strided / dilated blocks are not covered")
        self.use_conv_mlp = False
        self.conv_dw = nn.Conv2d(in_chs, out_chs, kernel_size, stride=1, padding=kernel_size // 2, groups=in_chs,
                                 bias=conv_bias, **dd)
        self.norm = (norm_layer or nn.LayerNorm)(out_chs, **dd)
        self.mlp = (GlobalResponseNormMlp if use_grn else Mlp)(out_chs, int(mlp_ratio * out_chs), act_layer=act_layer, **dd)
        self.gamma = nn.Parameter(ls_init_value * torch.ones(out_chs, **dd)) if ls_init_value is not None else None
        self.shortcut = nn.Identity()
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()

    def forward_nhwc(self, x, B, H, W):
        return block_forward_nhwc(self, x, B, H, W)

    def forward(self, x):
        B, C, H, W = x.shape
        rows = x.permute(0, 2, 3, 1).contiguous().view(B * H * W, C)
        return self.forward_nhwc(rows, B, H, W).view(B, H, W, C).permute(0, 3, 1, 2)


def _mlp_w(lin):
    """GEMM operand of an Mlp layer: nn.Linear weight, or the [out, in, 1, 1] weight of the 1x1 Conv2d a `conv_mlp=True`
    block uses (ConvMlp, timm/layers/mlp.py; convnext.py:176-181) viewed [out, in] -- on channels-last rows the two are the
    same matmul."""
    w = _w(lin)
    return w.reshape(w.shape[0], -1) if w.dim() == 4 else w


def _check_block(blk):
    """Raise B200Error unless `blk` (this mirror's or timm's ConvNeXtBlock) is the configuration the kernels cover."""
    conv = blk.conv_dw
    k = conv.kernel_size[0]
    conv_mlp = getattr(blk, "use_conv_mlp", False)
    if conv_mlp:      # LayerNorm2d + 1x1-conv MLP (ConvNeXt-V2 atto ... nano): per-row LayerNorm + GEMMs on channels-last rows
        fc1 = blk.mlp.fc1
        ok = (type(blk.norm).__name__ == "LayerNorm2d" and isinstance(fc1, nn.Conv2d) and fc1.kernel_size == (1, 1)
              and fc1.stride == (1, 1) and fc1.groups == 1 and blk.mlp.fc2.kernel_size == (1, 1)
              and isinstance(getattr(blk.mlp, "norm", nn.Identity()), nn.Identity))
    else:
        ok = type(blk.norm).__name__ == "LayerNorm"
    if (not ok or not isinstance(blk.shortcut, nn.Identity) or conv.stride != (1, 1)
            or conv.dilation != (1, 1) or conv.groups != conv.in_channels or conv.in_channels != conv.out_channels
            or conv.kernel_size != (k, k) or conv.padding != (k // 2, k // 2)):
        raise B200Error("ConvNeXtBlock: strided / dilated blocks and non-1x1 conv MLPs are not covered")
    if mlp_act(blk.mlp) != 0 and hasattr(blk.mlp, "grn"):
        raise B200Error("GlobalResponseNormMlp: only the exact GELU is covered")


def block_forward_nhwc(self, x, B, H, W):
    """x: [B*H*W, C] bf16 rows (channels-last).  convnext.py:197-213 without the permutes.  `self` is duck-typed:
    this mirror's ConvNeXtBlock or timm's own (the plugin rebinds timm's forward onto this function)."""
    C = x.shape[1]
    _check_block(self)
    # DropPath (convnext.py:213, drop.py:158-175): the per-sample factor is drawn exactly as the reference draws it --
    # bernoulli_(keep_prob) on B elements of the CUDA generator -- and applied inside the fc2 epilogue
    rs = None
    dp = getattr(self, "drop_path", None)
    if self.training and getattr(dp, "drop_prob", 0.) > 0:
        keep = 1.0 - dp.drop_prob
        rs = x.new_empty((B, 1, 1, 1)).bernoulli_(keep).reshape(-1).float()
        if keep > 0.0 and getattr(dp, "scale_by_keep", True):
            rs = rs / keep
        rs = rs.contiguous()
    grn = getattr(self.mlp, "grn", None)     # GlobalResponseNormMlp (ConvNeXt-V2, mlp.py:251-295)
    if x.requires_grad or _needs_grad(self):
        extra = (None, None, 1e-6) if grn is None else (grn.weight, grn.bias, grn.eps)
        return _BlockFn.apply(x, _w(self.conv_dw), self.conv_dw.bias, self.norm.weight, self.norm.bias,
                              _mlp_w(self.mlp.fc1), self.mlp.fc1.bias, _mlp_w(self.mlp.fc2), self.mlp.fc2.bias,
                              self.gamma, rs, B, H, W, self.norm.eps, *extra, mlp_act(self.mlp))
    # the block's LayerNorm folded into its neighbours (inference): the depthwise convolution leaves per-pixel partial sums,
    # fc1 runs on the un-normalised rows against weight * gamma and finishes LN in its epilogue (functional._ln_fold)
    wt = _dw_weight_t(self.conv_dw)
    act = mlp_act(self.mlp)
    fold = None
    if B * H * W >= 256 and act == 0 and ops.dwconv_ln_fusable(C, wt.shape[0]) and x.data_ptr() % 16 == 0:
        fold = BF._ln_fold(self.norm, self.mlp.fc1, _mlp_w(self.mlp.fc1), self.mlp.fc1.bias, "fc1")
    if fold is not None:
        y, st = ops.dwconv_nhwc_ln(x.view(B, H, W, C), wt, self.conv_dw.bias)
        a = ops.linear_fwd(y.view(B * H * W, C), fold[0], fold[2], epilogue=ops.EPI_BIAS_GELU, ln_in=(st, fold[1], fold[3]))
    else:
        y = ops.dwconv_nhwc(x.view(B, H, W, C), wt, self.conv_dw.bias).view(B * H * W, C)
        h = ops.layernorm_fwd(y, self.norm.weight, self.norm.bias, self.norm.eps)[0]
        a = ops.linear_fwd(h, _mlp_w(self.mlp.fc1), self.mlp.fc1.bias, epilogue=ops.EPI_BIAS_GELU, act=act)
    if grn is not None:
        ops.grn_fwd(a, grn.weight, grn.bias, B, H * W, grn.eps)
    return ops.linear_fwd(a, _mlp_w(self.mlp.fc2), self.mlp.fc2.bias, epilogue=ops.EPI_BIAS_RESIDUAL, aux=x,
                          gamma=self.gamma, row_scale=rs, rows_per_scale=H * W)


class ConvNeXtStage(nn.Module):
    def __init__(self, in_chs, out_chs, kernel_size=7, stride=2, depth=2, dilation=(1, 1), drop_path_rates=None,
                 ls_init_value=1.0, conv_mlp=False, conv_bias=True, use_grn=False, act_layer=nn.GELU, norm_layer=None,
                 norm_layer_cl=None, device=None, dtype=None):
        super().__init__()
        dd = {"device": device, "dtype": dtype}
        self.grad_checkpointing = False
        if in_chs != out_chs or stride > 1:
            ds_ks = 2 if stride > 1 else 1
            self.downsample = nn.Sequential(
                (norm_layer or LayerNorm2d)(in_chs, **dd),
                nn.Conv2d(in_chs, out_chs, kernel_size=ds_ks, stride=stride, bias=conv_bias, **dd))
            in_chs = out_chs
        else:
            self.downsample = nn.Identity()
        drop_path_rates = drop_path_rates or [0.] * depth
        self.blocks = nn.Sequential(*[
            ConvNeXtBlock(in_chs=out_chs, out_chs=out_chs, kernel_size=kernel_size, drop_path=drop_path_rates[i],
                          ls_init_value=ls_init_value, conv_bias=conv_bias, use_grn=use_grn, act_layer=act_layer,
                          norm_layer=norm_layer_cl, **dd)
            for i in range(depth)])

    def forward_nhwc(self, x, B, H, W):
        return stage_forward_nhwc(self, x, B, H, W)


def stage_forward_nhwc(self, x, B, H, W):
    """ConvNeXtStage.forward (convnext.py:300-309) on channels-last rows; duck-typed like block_forward_nhwc."""
    if not isinstance(self.downsample, nn.Identity):
        norm, conv = self.downsample[0], self.downsample[1]
        p = conv.kernel_size[0]
        C = x.shape[1]
        if (len(self.downsample) != 2 or not isinstance(conv, nn.Conv2d) or conv.kernel_size != (p, p)
                or conv.stride != (p, p) or conv.padding != (0, 0) or conv.groups != 1 or p not in (1, 2)
                or type(norm).__name__ != "LayerNorm2d" or H % p or W % p):
            raise B200Error("ConvNeXtStage: only LayerNorm2d + 1x1 / 2x2-stride-2 conv downsampling is covered")
        if x.requires_grad or _needs_grad(self.downsample):
            x = _DownsampleFn.apply(x, norm.weight, norm.bias, _w(conv), conv.bias, B, H, W, norm.eps)
            H, W = H // p, W // p
        else:
            h = ops.layernorm_fwd(x, norm.weight, norm.bias, norm.eps)[0]
            w2 = _conv_weight_khwc(conv)
            if p > 1:
                h = ops.patchify_nhwc(h.view(B, H, W, C), p)
                H, W = H // p, W // p
            x = ops.linear_fwd(h, w2, conv.bias)
    if getattr(self, "grad_checkpointing", False) and torch.is_grad_enabled() and x.requires_grad:
        # checkpoint_seq(self.blocks, x) (convnext.py:305-306, _manipulate.py:213-286): a block's activations are dropped
        # after its forward and its kernels run again in the backward (non-reentrant checkpoint: the CUDA RNG state is
        # restored first, so a DropPath draw repeats)
        from torch.utils.checkpoint import checkpoint
        for blk in self.blocks:
            x = checkpoint(block_forward_nhwc, blk, x, B, H, W, use_reentrant=False)
    else:
        for blk in self.blocks:
            x = block_forward_nhwc(blk, x, B, H, W)
    return x, H, W


def _dw_weight_t(conv):
    """Depthwise Conv2d weight [C,1,k,k] -> cached [k,k,C] (the layout the dwconv kernel reads 16 bytes per tap)."""
    w = conv.weight
    key = ops.wkey(w)
    cache = getattr(conv, "_b200_wt", None)
    if cache is None or cache[0] != key:
        conv._b200_wt = (key, ops.dwconv_weight_t(w.detach().to(torch.bfloat16)))
        cache = conv._b200_wt
    return cache[1]


def _conv_weight_khwc(conv):
    """Conv2d weight [Cout, Cin, kh, kw] -> cached contiguous [Cout, kh*kw*Cin] matching patchify_nhwc's k order."""
    w = conv.weight
    key = ops.wkey(w)
    cache = getattr(conv, "_b200_khwc", None)
    if cache is None or cache[0] != key:
        conv._b200_khwc = (key, w.detach().to(torch.bfloat16).permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous())
        cache = conv._b200_khwc
    return cache[1]


class NormMlpClassifierHead(nn.Module):
    """timm.layers.NormMlpClassifierHead (classifier.py:145-220) without the optional hidden layer."""

    def __init__(self, in_features, num_classes, hidden_size=None, pool_type="avg", drop_rate=0., norm_layer=None,
                 act_layer=None, device=None, dtype=None):
        super().__init__()
        dd = {"device": device, "dtype": dtype}
        if hidden_size or pool_type != "avg":
            raise B200Error("NormMlpClassifierHead: hidden layer / non-avg pooling are not covered")
        self.in_features = self.num_features = in_features
        self.norm = (norm_layer or LayerNorm2d)(in_features, **dd)
        self.drop = nn.Dropout(drop_rate)
        self.fc = nn.Linear(in_features, num_classes, **dd) if num_classes > 0 else nn.Identity()

    def forward_nhwc(self, x, B, HW, pre_logits=False):
        return head_forward_nhwc(self, x, B, HW, pre_logits)


def head_forward_nhwc(self, x, B, HW, pre_logits=False):
    """NormMlpClassifierHead.forward (classifier.py:211-220): avg-pool -> LayerNorm2d -> fc, on channels-last rows."""
    pool = getattr(self, "global_pool", None)
    if ((pool is not None and getattr(pool, "pool_type", "avg") != "avg") or type(self.norm).__name__ != "LayerNorm2d"
            or not isinstance(getattr(self, "pre_logits", nn.Identity()), nn.Identity)
            or (self.training and getattr(self.drop, "p", 0.) > 0)):
        raise B200Error("NormMlpClassifierHead: hidden layer / non-avg pooling / head dropout are not covered")
    if x.requires_grad or _needs_grad(self):
        from .autograd import LayerNormFn, LinearFn
        h = LayerNormFn.apply(_AvgPoolFn.apply(x, B, HW), self.norm.weight, self.norm.bias, self.norm.eps)
        if pre_logits or isinstance(self.fc, nn.Identity):
            return h
        return LinearFn.apply(h, _w(self.fc), self.fc.bias)
    pooled = ops.avgpool_nhwc(x.view(B, HW, -1))
    h = ops.layernorm_fwd(pooled, self.norm.weight, self.norm.bias, self.norm.eps)[0]
    if pre_logits or isinstance(self.fc, nn.Identity):
        return h
    from .functional import _padded_out_linear
    if self.fc.weight.shape[0] % 8:
        return _padded_out_linear(self.fc, h)
    return ops.linear_fwd(h, _w(self.fc), self.fc.bias)


class ConvNeXt(nn.Module):
    """timm.models.ConvNeXt (convnext.py:338-629), patch stem / LayerNorm / Linear-MLP configuration."""

    def __init__(self, in_chans: int = 3, num_classes: int = 1000, global_pool: str = "avg", output_stride: int = 32,
                 depths: Tuple[int, ...] = (3, 3, 9, 3), dims: Tuple[int, ...] = (96, 192, 384, 768), kernel_sizes=7,
                 ls_init_value: Optional[float] = 1e-6, stem_type: str = "patch", patch_size: int = 4,
                 head_init_scale: float = 1., head_norm_first: bool = False, head_hidden_size=None,
                 conv_mlp: bool = False, conv_bias: bool = True, use_grn: bool = False, act_layer=None,
                 norm_layer=None, norm_eps: Optional[float] = None, drop_rate: float = 0., drop_path_rate: float = 0.,
                 device=None, dtype=None):
        super().__init__()
        dd = {"device": device, "dtype": dtype}
        acts = {None: nn.GELU, "gelu": nn.GELU, nn.GELU: nn.GELU, "silu": nn.SiLU, nn.SiLU: nn.SiLU,
                "gelu_tanh": lambda: nn.GELU(approximate="tanh")}       # create_act.py keys timm's ConvNeXt configs use
        if (stem_type != "patch" or output_stride != 32 or head_norm_first or conv_mlp or norm_layer
                or global_pool != "avg" or act_layer not in acts):
            raise B200Error("ConvNeXt: only the patch-stem / LayerNorm / Linear-MLP / GELU, tanh-GELU, SiLU / avg-pool "
                            "configuration is covered by the B200 path")
        act_cls = acts[act_layer]
        eps = norm_eps if norm_eps is not None else 1e-6
        ks = kernel_sizes if isinstance(kernel_sizes, (tuple, list)) else (kernel_sizes,) * 4
        ln2d = lambda c, **kw: LayerNorm2d(c, eps=eps, **kw)        # noqa: E731
        ln = lambda c, **kw: nn.LayerNorm(c, eps=eps, **kw)         # noqa: E731
        self.num_classes, self.in_chans, self.drop_rate = num_classes, in_chans, drop_rate
        self.stem = nn.Sequential(nn.Conv2d(in_chans, dims[0], kernel_size=patch_size, stride=patch_size,
                                            bias=conv_bias, **dd), ln2d(dims[0], **dd))
        total = sum(depths)
        rates = [x.item() for x in torch.linspace(0, drop_path_rate, total)]
        stages, prev, k = [], dims[0], 0
        for i in range(4):
            stages.append(ConvNeXtStage(prev, dims[i], kernel_size=ks[i], stride=2 if i > 0 else 1, depth=depths[i],
                                        drop_path_rates=rates[k:k + depths[i]], ls_init_value=ls_init_value,
                                        conv_bias=conv_bias, use_grn=use_grn, act_layer=act_cls, norm_layer=ln2d, norm_layer_cl=ln, **dd))
            k += depths[i]
            prev = dims[i]
        self.stages = nn.Sequential(*stages)
        self.num_features = self.head_hidden_size = prev
        self.norm_pre = nn.Identity()
        self.head = NormMlpClassifierHead(prev, num_classes, pool_type=global_pool, drop_rate=drop_rate,
                                          norm_layer=ln2d, **dd)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        """timm/models/convnext.py _init_weights: trunc_normal(.02) conv / linear weights, zero biases."""
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            nn.init.trunc_normal_(m.weight, std=.02, a=-2., b=2.)
            if m.bias is not None:
                nn.init.zeros_(m.bias)

    def get_classifier(self):
        return self.head.fc

    def _features_nhwc(self, img):
        return features_nhwc(self, img)

    def forward_features(self, img):
        """NCHW feature map like timm's forward_features (convnext.py:606-611); the permute is a view."""
        x, B, H, W = self._features_nhwc(img)
        return x.view(B, H, W, -1).permute(0, 3, 1, 2)

    def forward_head(self, x, pre_logits: bool = False):
        B, C, H, W = x.shape
        rows = x.permute(0, 2, 3, 1).contiguous().view(B * H * W, C)
        return self.head.forward_nhwc(rows, B, H * W, pre_logits)

    def forward(self, img):
        x, B, H, W = self._features_nhwc(img)
        return self.head.forward_nhwc(x, B, H * W)


def features_nhwc(self, img):
    """stem -> stages (convnext.py:606-611) on channels-last rows; `self` is this mirror's ConvNeXt or timm's."""
    if not img.is_cuda:
        raise B200Error("libb200timm kernels need CUDA tensors (there is no CPU path)")
    if img.dtype not in (torch.bfloat16, torch.float32, torch.uint8):
        raise B200Error("ConvNeXt B200 path needs bf16 images (fp32 under autocast with fp32 master weights; uint8 "
                        "loader batches with ops.set_input_norm)")
    if (len(self.stem) != 2 or not isinstance(self.stem[0], nn.Conv2d) or type(self.stem[1]).__name__ != "LayerNorm2d"
            or not isinstance(getattr(self, "norm_pre", nn.Identity()), nn.Identity)):
        raise B200Error("ConvNeXt: only the patch stem (conv + LayerNorm2d) without norm_pre is covered")
    conv, norm = self.stem[0], self.stem[1]
    p = conv.kernel_size[0]
    if conv.kernel_size != (p, p) or conv.stride != (p, p) or conv.padding != (0, 0):
        raise B200Error("ConvNeXt: the stem must be a non-overlapping patch conv")
    B, C, H, W = img.shape
    if H % p or W % p:
        raise B200Error("image size must be divisible by the stem patch size")
    K = C * p * p
    kpad = (K + 7) // 8 * 8
    wc = _w(conv)
    w2 = wc.view(wc.shape[0], K)
    if kpad != K:
        raise B200Error("stem patch size with C*p*p not a multiple of 8 is not covered")
    if _needs_grad(self.stem):
        from .autograd import LayerNormFn, PatchEmbedFn
        x = PatchEmbedFn.apply(img, wc, conv.bias, p).reshape(B * (H // p) * (W // p), -1)
        x = LayerNormFn.apply(x, norm.weight, norm.bias, norm.eps)
    else:
        x = ops.gemm(ops.patchify(img.contiguous(), p, kpad), w2, B * (H // p) * (W // p), w2.shape[0], kpad,
                     bias=conv.bias)
        x = ops.layernorm_fwd(x, norm.weight, norm.bias, norm.eps)[0]
    H, W = H // p, W // p
    for stage in self.stages:
        x, H, W = stage_forward_nhwc(stage, x, B, H, W)
    return x, B, H, W


MODEL_ARGS = {
    "convnext_tiny": dict(depths=(3, 3, 9, 3), dims=(96, 192, 384, 768)),
    "convnext_small": dict(depths=(3, 3, 27, 3), dims=(96, 192, 384, 768)),
    "convnext_base": dict(depths=(3, 3, 27, 3), dims=(128, 256, 512, 1024)),       # convnext.py:1293-1296
    "convnext_large": dict(depths=(3, 3, 27, 3), dims=(192, 384, 768, 1536)),
    # ConvNeXt-V2 (Global Response Norm in the MLP, no LayerScale), convnext.py:1365-1387
    "convnextv2_tiny": dict(depths=(3, 3, 9, 3), dims=(96, 192, 384, 768), use_grn=True, ls_init_value=None),
    "convnextv2_base": dict(depths=(3, 3, 27, 3), dims=(128, 256, 512, 1024), use_grn=True, ls_init_value=None),
    "convnextv2_large": dict(depths=(3, 3, 27, 3), dims=(192, 384, 768, 1536), use_grn=True, ls_init_value=None),
}


def create_model(name: str, **kwargs) -> ConvNeXt:
    if name not in MODEL_ARGS:
        raise B200Error(f"unknown ConvNeXt variant {name!r}; mirrored: {sorted(MODEL_ARGS)}")
    return ConvNeXt(**dict(MODEL_ARGS[name], **kwargs))
