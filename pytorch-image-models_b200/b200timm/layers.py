"""Drop-in mirrors of timm's hot-path layer modules: same constructor arguments, same parameter / sub-module
names (hence identical ``state_dict()`` keys), forwards running on the libb200timm kernels.

Reference classes (timm 1.0.28.dev0): timm/layers/attention.py:43 (Attention), mlp.py:14 (Mlp),
norm.py:70 (LayerNorm), layer_scale.py:5 (LayerScale), drop.py:178 (DropPath), patch_embed.py:26 (PatchEmbed).
The Linear / Conv2d / LayerNorm children stay real torch parameter holders (timm's weight init, checkpoint
loaders and optimizers reach into them by name) -- only ``forward`` differs.
"""
from typing import Optional, Type

import torch
from torch import nn

from . import functional as BF
from ._lib import B200Error


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class LayerNorm(nn.LayerNorm):
    """timm.layers.LayerNorm: nn.LayerNorm with eps=1e-6 default (norm.py:70-90)."""

    def __init__(self, num_channels: int, eps: float = 1e-6, affine: bool = True, device=None, dtype=None):
        super().__init__(num_channels, eps=eps, elementwise_affine=affine, device=device, dtype=dtype)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return BF.layernorm(self, x)


class LayerScale(nn.Module):
    """timm.layers.LayerScale (layer_scale.py:5-27): x * gamma.  Inside a Block it is fused into the
    preceding GEMM's epilogue; standalone it is a (rare) elementwise op and is not provided as a kernel."""

    def __init__(self, dim: int, init_values: float = 1e-5, inplace: bool = False, device=None, dtype=None):
        super().__init__()
        self.inplace = inplace
        self.gamma = nn.Parameter(init_values * torch.ones(dim, device=device, dtype=dtype))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        raise B200Error("LayerScale is fused into the proj / fc2 GEMM epilogue by Block.forward; "
                        "calling it standalone is not on the B200 hot path")


class DropPath(nn.Module):
    """timm.layers.DropPath (drop.py:178-190): per-sample stochastic depth.  Identity in eval mode; in
    training the per-sample keep mask is fused into the residual GEMM epilogue by Block.forward."""

    def __init__(self, drop_prob: float = 0., scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0. or not self.training:
            return x
        raise B200Error("DropPath in training mode is fused into Block.forward; standalone use is not supported")

    def extra_repr(self):
        return f"drop_prob={round(self.drop_prob, 3):0.3f}"


class Mlp(nn.Module):
    """timm.layers.Mlp (mlp.py:14-54)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, norm_layer=None,
                 bias=True, drop=0., use_conv=False, device=None, dtype=None):
        super().__init__()
        if use_conv:
            raise B200Error("Mlp(use_conv=True) is not supported")
        dd = {"device": device, "dtype": dtype}
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        bias = _pair(bias)
        drop_probs = _pair(drop)
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias[0], **dd)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop_probs[0])
        self.norm = norm_layer(hidden_features, **dd) if norm_layer is not None else nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias[1], **dd)
        self.drop2 = nn.Dropout(drop_probs[1])

    def forward(self, x):
        return BF.mlp_forward(self, x)


class GluMlp(nn.Module):
    """timm.layers.GluMlp (mlp.py:57-113): fc1 of width `hidden_features`, split into value / gate halves."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.Sigmoid, norm_layer=None,
                 bias=True, drop=0., use_conv=False, gate_last=True, device=None, dtype=None):
        super().__init__()
        if use_conv:
            raise B200Error("GluMlp(use_conv=True) is not supported")
        dd = {"device": device, "dtype": dtype}
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        assert hidden_features % 2 == 0
        bias = _pair(bias)
        drop_probs = _pair(drop)
        self.chunk_dim = -1
        self.gate_last = gate_last
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias[0], **dd)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop_probs[0])
        self.norm = norm_layer(hidden_features // 2, **dd) if norm_layer is not None else nn.Identity()
        self.fc2 = nn.Linear(hidden_features // 2, out_features, bias=bias[1], **dd)
        self.drop2 = nn.Dropout(drop_probs[1])

    def init_weights(self):
        if self.fc1.bias is not None:
            nn.init.ones_(self.fc1.bias[self.fc1.bias.shape[0] // 2:])
        nn.init.normal_(self.fc1.weight[self.fc1.weight.shape[0] // 2:], std=1e-6)

    def forward(self, x):
        return BF.glu_mlp_forward(self, x)


def SwiGLUPacked(*args, **kwargs):
    """timm.layers.SwiGLUPacked (mlp.py:116): GluMlp with SiLU on the FIRST half."""
    kwargs.setdefault("act_layer", nn.SiLU)
    kwargs.setdefault("gate_last", False)
    return GluMlp(*args, **kwargs)


class SwiGLU(nn.Module):
    """timm.layers.SwiGLU (mlp.py:119-162): separate gate / value projections fc1_g, fc1_x."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.SiLU, norm_layer=None,
                 bias=True, drop=0., align_to=0, device=None, dtype=None):
        super().__init__()
        dd = {"device": device, "dtype": dtype}
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        bias = _pair(bias)
        drop_probs = _pair(drop)
        if align_to:
            hidden_features = hidden_features + (-hidden_features % align_to)
        self.fc1_g = nn.Linear(in_features, hidden_features, bias=bias[0], **dd)
        self.fc1_x = nn.Linear(in_features, hidden_features, bias=bias[0], **dd)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop_probs[0])
        self.norm = norm_layer(hidden_features, **dd) if norm_layer is not None else nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias[1], **dd)
        self.drop2 = nn.Dropout(drop_probs[1])

    def init_weights(self):
        if self.fc1_g.bias is not None:
            nn.init.ones_(self.fc1_g.bias)
        nn.init.normal_(self.fc1_g.weight, std=1e-6)

    def forward(self, x):
        return BF.glu_mlp_forward(self, x)


class AttentionPoolLatent(nn.Module):
    """timm.layers.AttentionPoolLatent (attention_pool.py:13-128), the 'map' pooling head of SigLIP-style ViTs: one
    learned latent query attends over the tokens, followed by proj and x + mlp(norm(x))."""

    def __init__(self, in_features: int, out_features: int = None, embed_dim: int = None, num_heads: int = 8,
                 feat_size: Optional[int] = None, mlp_ratio: float = 4.0, qkv_bias: bool = True, qk_norm: bool = False,
                 latent_len: int = 1, latent_dim: int = None, pos_embed: str = "", pool_type: str = "token",
                 norm_layer=None, act_layer=nn.GELU, drop: float = 0.0, device=None, dtype=None):
        super().__init__()
        dd = {"device": device, "dtype": dtype}
        embed_dim = embed_dim or in_features
        if out_features is None:
            out_features = in_features
        if qk_norm or latent_len != 1 or pos_embed or pool_type != "token" or embed_dim != in_features:
            raise B200Error("AttentionPoolLatent: qk_norm / latent_len > 1 / position embedding / avg pooling are not covered")
        assert embed_dim % num_heads == 0
        self.num_heads, self.head_dim = num_heads, embed_dim // num_heads
        self.feat_size, self.scale, self.pool = feat_size, (embed_dim // num_heads) ** -0.5, pool_type
        self.pos_embed = None
        self.latent_dim, self.latent_len = latent_dim or embed_dim, latent_len
        self.latent = nn.Parameter(torch.zeros(1, latent_len, embed_dim, **dd))
        self.q = nn.Linear(embed_dim, embed_dim, bias=qkv_bias, **dd)
        self.kv = nn.Linear(embed_dim, embed_dim * 2, bias=qkv_bias, **dd)
        self.q_norm, self.k_norm = nn.Identity(), nn.Identity()
        if out_features > 0:
            self.proj = nn.Linear(embed_dim, out_features, **dd)
            self.proj_drop = nn.Dropout(drop)
            self.norm = norm_layer(out_features, **dd) if norm_layer is not None else nn.Identity()
            self.mlp = Mlp(out_features, int(out_features * mlp_ratio), out_features=out_features, act_layer=act_layer, **dd)
        else:
            self.proj, self.proj_drop, self.norm, self.mlp = nn.Identity(), nn.Dropout(drop), nn.Identity(), None
            out_features = embed_dim
        self.out_features = out_features
        nn.init.trunc_normal_(self.latent, std=self.latent_dim ** -0.5)

    def forward(self, x, attn_mask: Optional[torch.Tensor] = None):
        if attn_mask is not None:
            raise B200Error("AttentionPoolLatent: masks are not covered")
        return BF.attention_pool_forward(self, x)


class RmsNorm(nn.Module):
    """timm.layers.RmsNorm (norm.py:202-248): y = x * rsqrt(mean(x^2) + eps) * weight."""

    def __init__(self, channels: int, eps: float = 1e-6, affine: bool = True, device=None, dtype=None):
        super().__init__()
        self.normalized_shape = (channels,)
        self.eps = eps
        self.elementwise_affine = affine
        if affine:
            self.weight = nn.Parameter(torch.ones(channels, device=device, dtype=dtype))
        else:
            self.register_parameter("weight", None)

    def forward(self, x):
        return BF.layernorm(self, x)


class Attention(nn.Module):
    """timm.layers.Attention (attention.py:43-139)."""

    def __init__(self, dim: int, num_heads: int = 8, attn_head_dim: Optional[int] = None,
                 dim_out: Optional[int] = None, qkv_bias: bool = False, qk_norm: bool = False,
                 scale_norm: bool = False, proj_bias: bool = True, attn_drop: float = 0., proj_drop: float = 0.,
                 norm_layer: Optional[Type[nn.Module]] = None, device=None, dtype=None):
        super().__init__()
        dd = {"device": device, "dtype": dtype}
        dim_out = dim_out or dim
        head_dim = attn_head_dim
        if head_dim is None:
            assert dim % num_heads == 0, "dim should be divisible by num_heads"
            head_dim = dim // num_heads
        if qk_norm or scale_norm:
            assert norm_layer is not None, "norm_layer must be provided if qk_norm or scale_norm is True"
        self.num_heads = num_heads
        self.head_dim = head_dim
        self.attn_dim = num_heads * head_dim
        self.scale = head_dim ** -0.5
        self.fused_attn = True
        self.qkv = nn.Linear(dim, self.attn_dim * 3, bias=qkv_bias, **dd)
        self.q_norm = norm_layer(head_dim, **dd) if qk_norm else nn.Identity()
        self.k_norm = norm_layer(head_dim, **dd) if qk_norm else nn.Identity()
        self.attn_drop = nn.Dropout(attn_drop)
        self.norm = norm_layer(self.attn_dim, **dd) if scale_norm else nn.Identity()
        self.proj = nn.Linear(self.attn_dim, dim_out, bias=proj_bias, **dd)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x: torch.Tensor, attn_mask: Optional[torch.Tensor] = None, is_causal: bool = False,
                kv_len: Optional[torch.Tensor] = None) -> torch.Tensor:
        if kv_len is None:
            kv_len = attn_mask_operand(attn_mask, is_causal, x.shape[1], x.shape[0], x.device)
        return BF.attention_forward(self, x, kv_len)


class AttentionRope(nn.Module):
    """timm.layers.AttentionRope (attention.py:143-282), fused-QKV configuration: the rotary embedding is applied to
    q and k inside the QKV projection's epilogue (interleaved layout) or by one in-place pass (half layout)."""

    def __init__(self, dim: int, num_heads: int = 8, dim_out: Optional[int] = None, qkv_bias: bool = True,
                 qkv_fused: bool = True, num_prefix_tokens: int = 1, attn_drop: float = 0., proj_drop: float = 0.,
                 attn_head_dim: Optional[int] = None, norm_layer: Optional[Type[nn.Module]] = None,
                 qk_norm: bool = False, scale_norm: bool = False, proj_bias: bool = True, rotate_half: bool = False,
                 device=None, dtype=None):
        super().__init__()
        dd = {"device": device, "dtype": dtype}
        if not qkv_fused or qk_norm or scale_norm:
            raise B200Error("AttentionRope: separate q/k/v projections, qk-norm and scale-norm are not covered")
        dim_out = dim_out or dim
        head_dim = attn_head_dim
        if head_dim is None:
            assert dim % num_heads == 0, "dim should be divisible by num_heads"
            head_dim = dim // num_heads
        self.num_heads = num_heads
        self.head_dim = head_dim
        self.attn_dim = head_dim * num_heads
        self.scale = head_dim ** -0.5
        self.num_prefix_tokens = num_prefix_tokens
        self.fused_attn = True
        self.rotate_half = rotate_half
        self.qkv = nn.Linear(dim, self.attn_dim * 3, bias=qkv_bias, **dd)
        self.q_proj = self.k_proj = self.v_proj = None
        self.q_norm = nn.Identity()
        self.k_norm = nn.Identity()
        self.attn_drop = nn.Dropout(attn_drop)
        self.norm = nn.Identity()
        self.proj = nn.Linear(self.attn_dim, dim_out, bias=proj_bias, **dd)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x: torch.Tensor, rope: Optional[torch.Tensor] = None, attn_mask: Optional[torch.Tensor] = None,
                is_causal: bool = False, kv_len: Optional[torch.Tensor] = None) -> torch.Tensor:
        if kv_len is None:
            kv_len = attn_mask_operand(attn_mask, is_causal, x.shape[1], x.shape[0], x.device)
        return BF.attention_rope_forward(self, x, rope, kv_len)


_KV_CACHE = {}      # id(mask) -> (weakref to the mask, version, n_tokens, batch, operand): one conversion per forward, not per block
_CAUSAL_CACHE = {}  # (n_tokens, device) -> fp32 [1, 1, N, ld] causal bias


def _as_key_padding(m: torch.Tensor, n_tokens: int):
    """int32 [B] valid-key counts when the bool keep-mask `m` ([Bm, Hm, Nq or 1, N]) is a pure prefix-valid key-padding mask:
    every query row that is valid itself keeps exactly the first kv_len keys, for every head (NaFlex:
    naflexvit.py:964-1029 builds q_valid & k_valid; the rows of padded queries do not matter).  None otherwise."""
    first = m[:, 0]                                                   # [Bm, Nq or 1, N]
    keys = first.any(dim=-2)                                          # key kept by at least one query
    kv_len = keys.sum(-1).to(torch.int32)
    ar = torch.arange(n_tokens, device=m.device)
    expect = ar[None, :] < kv_len[:, None]                            # [Bm, N]
    ok = (keys == expect).all()
    if first.shape[-2] != 1:
        rows_valid = ar[None, :, None] < kv_len[:, None, None]        # query rows inside the valid prefix
        ok = ok & ((first == expect[:, None, :]) | ~rows_valid).all()
    if m.shape[1] != 1:
        ok = ok & (m == first[:, None]).all()
    return kv_len if bool(ok) else None                               # one host read, cached by the caller


def attn_bias_operand(attn_mask: torch.Tensor, n_tokens: int) -> torch.Tensor:
    """The additive fp32 bias libb200timm's b200_attn_fwd_bias / _bwd_bias take, from an attention mask as
    resolve_self_attn_mask (timm/layers/attention.py:21-40) interprets it: bool = keep-mask -> 0 / -inf, float = additive.
    Result: contiguous fp32 [Bm, Hm, N, ld] (ld = N rounded up to 4; Bm / Hm = the mask's own batch / head extent, 1 = broadcast)."""
    m = attn_mask
    while m.dim() < 4:
        m = m.unsqueeze(0)
    if m.dim() != 4 or m.shape[-1] != n_tokens or m.shape[-2] not in (1, n_tokens):
        raise B200Error(f"attn_mask of shape {tuple(attn_mask.shape)} does not broadcast to [B, H, {n_tokens}, {n_tokens}]")
    ld = (n_tokens + 3) // 4 * 4
    out = torch.zeros(m.shape[0], m.shape[1], n_tokens, ld, dtype=torch.float32, device=m.device)
    if m.dtype == torch.bool:
        out[..., :n_tokens].masked_fill_(~m.expand(-1, -1, n_tokens, -1), float("-inf"))
    else:
        out[..., :n_tokens] = m.expand(-1, -1, n_tokens, -1).float()
    return out


def causal_bias_operand(n_tokens: int, device) -> torch.Tensor:
    """-inf above the diagonal (attention.py:31-32), fp32 [1, 1, N, ld]; built once per (N, device)."""
    key = (n_tokens, str(device))
    hit = _CAUSAL_CACHE.get(key)
    if hit is None:
        ld = (n_tokens + 3) // 4 * 4
        hit = torch.zeros(1, 1, n_tokens, ld, dtype=torch.float32, device=device)
        hit[0, 0, :, :n_tokens] = torch.full((n_tokens, n_tokens), float("-inf"), device=device).triu_(1)
        _CAUSAL_CACHE[key] = hit
    return hit


def attn_mask_operand(attn_mask: Optional[torch.Tensor], is_causal: bool, n_tokens: int, batch: Optional[int] = None,
                      device=None, strict_prefix: bool = False):
    """What the attention kernels take in place of (attn_mask, is_causal) -- passed down as `kv_len`:
      None                       no mask
      int32 [B]                  a prefix-valid key-padding mask as valid-key counts (the NaFlex layout; b200_attn_fwd's kv_len)
      fp32 [Bm, Hm, N, ld]       anything else -- causal (which takes precedence over attn_mask, attention.py:30-32), general
                                 bool / additive masks -- as an additive bias (b200_attn_fwd_bias)
    The classification reads one flag back from the device, so the result is cached on the mask object: the blocks of a
    forward that receive the same mask tensor pay for one conversion."""
    import weakref
    if is_causal:
        return causal_bias_operand(n_tokens, device if device is not None else attn_mask.device)
    if attn_mask is None:
        return None
    key = id(attn_mask)
    hit = _KV_CACHE.get(key)
    if hit is not None and hit[0]() is attn_mask and hit[1:4] == (attn_mask._version, n_tokens, batch):
        if strict_prefix and hit[4].dtype != torch.int32:
            raise B200Error("only prefix-valid key-padding masks are supported (NaFlex layout)")
        return hit[4]
    m = attn_mask
    while m.dim() < 4:
        m = m.unsqueeze(0)
    keep = m if m.dtype == torch.bool else m > -1.0      # additive key-padding masks hold 0 (keep) or finfo.min / -inf (drop)
    operand = None
    if m.dtype == torch.bool or bool(((m == 0) | ~keep).all()):
        operand = _as_key_padding(keep, n_tokens)
    if operand is not None:
        if batch is not None and operand.numel() == 1 and batch > 1:
            operand = operand.expand(batch)
        operand = operand.contiguous()
    elif strict_prefix:
        raise B200Error("only prefix-valid key-padding masks are supported (NaFlex layout)")
    else:
        operand = attn_bias_operand(attn_mask, n_tokens)
    if len(_KV_CACHE) > 64:
        for k in [k for k, v in _KV_CACHE.items() if v[0]() is None]:
            del _KV_CACHE[k]
    _KV_CACHE[key] = (weakref.ref(attn_mask), attn_mask._version, n_tokens, batch, operand)
    return operand


def BF_mask_to_kv_len(attn_mask: torch.Tensor, n_tokens: int, batch: Optional[int] = None) -> torch.Tensor:
    """Prefix-valid key-padding mask -> per-sample valid-key counts (int32 [B]); anything else raises (callers that can
    take a general mask use attn_mask_operand)."""
    return attn_mask_operand(attn_mask, False, n_tokens, batch, strict_prefix=True)


class PatchEmbed(nn.Module):
    """timm.layers.PatchEmbed (patch_embed.py:26-142), flatten=True / NLC output."""

    def __init__(self, img_size=224, patch_size: int = 16, in_chans: int = 3, embed_dim: int = 768,
                 norm_layer=None, flatten: bool = True, output_fmt: Optional[str] = None, bias: bool = True,
                 strict_img_size: bool = True, dynamic_img_pad: bool = False, device=None, dtype=None):
        super().__init__()
        if output_fmt is not None or not flatten:
            raise B200Error("PatchEmbed output formats other than flattened NLC are not supported")
        dd = {"device": device, "dtype": dtype}
        self.patch_size = _pair(patch_size)
        self.img_size, self.grid_size, self.num_patches = self._init_img_size(img_size)
        self.flatten = flatten
        self.strict_img_size = strict_img_size
        self.dynamic_img_pad = dynamic_img_pad
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias, **dd)
        self.norm = norm_layer(embed_dim, **dd) if norm_layer else nn.Identity()

    def _init_img_size(self, img_size):
        if img_size is None:
            return None, None, None
        img_size = _pair(img_size)
        grid_size = tuple(s // p for s, p in zip(img_size, self.patch_size))
        return img_size, grid_size, grid_size[0] * grid_size[1]

    def feat_ratio(self, as_scalar=True):
        return max(self.patch_size) if as_scalar else self.patch_size

    def dynamic_feat_size(self, img_size):
        return img_size[0] // self.patch_size[0], img_size[1] // self.patch_size[1]

    def forward(self, x):
        return BF.patch_embed_forward(self, x)
