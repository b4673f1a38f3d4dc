"""Mirror of timm's VisionTransformer for the hot-path configurations: same constructor keywords, same
module / parameter names (state_dict-compatible with timm checkpoints), forwards on libb200timm kernels.

Reference: timm/models/vision_transformer.py -- Block :128-214, VisionTransformer :711-1302,
init_weights_vit_timm :1305-1320, entrypoints :3105 (vit_tiny_patch16_224), :3188 (vit_base_patch16_224),
:3524 (vit_large_patch14_clip_336).  timm's factory / registry / pretrained machinery is NOT rebuilt here
(out of scope); `create_model` below only resolves the handful of names the benchmarks use.
"""
from functools import partial
from typing import Optional

import torch
from torch import nn

from . import functional as BF
from ._lib import B200Error
from .layers import Attention, DropPath, LayerNorm, LayerScale, Mlp, PatchEmbed


def calculate_drop_path_rates(drop_path_rate: float, depth: int):
    """timm/layers/drop.py:193-231 (per-block linear schedule)."""
    return [x.item() for x in torch.linspace(0, drop_path_rate, depth, device="cpu")]


class Block(nn.Module):
    """Pre-norm transformer block (vision_transformer.py:128-214)."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_norm=False, scale_attn_norm=False,
                 scale_mlp_norm=False, proj_bias=True, proj_drop=0., attn_drop=0., init_values=None, drop_path=0.,
                 act_layer=nn.GELU, norm_layer=LayerNorm, mlp_layer=Mlp, attn_layer=Attention, depth=0,
                 device=None, dtype=None):
        super().__init__()
        dd = {"device": device, "dtype": dtype}
        self.norm1 = norm_layer(dim, **dd)
        self.attn = attn_layer(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_norm=qk_norm,
                               scale_norm=scale_attn_norm, proj_bias=proj_bias, attn_drop=attn_drop,
                               proj_drop=proj_drop, norm_layer=norm_layer, **dd)
        self.ls1 = LayerScale(dim, init_values=init_values, **dd) if init_values else nn.Identity()
        self.drop_path1 = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim, **dd)
        self.mlp = mlp_layer(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer,
                             norm_layer=norm_layer if scale_mlp_norm else None, bias=proj_bias, drop=proj_drop, **dd)
        self.ls2 = LayerScale(dim, init_values=init_values, **dd) if init_values else nn.Identity()
        self.drop_path2 = DropPath(drop_path) if drop_path > 0. else nn.Identity()

    def forward(self, x, attn_mask: Optional[torch.Tensor] = None, is_causal: bool = False,
                kv_len: Optional[torch.Tensor] = None):
        if kv_len is None and (attn_mask is not None or is_causal):
            from .layers import attn_mask_operand
            kv_len = attn_mask_operand(attn_mask, is_causal, x.shape[1], x.shape[0], x.device)
        return BF.block_forward(self, x, kv_len)


def trunc_normal_(t, std=.02):
    """timm/layers/weight_init.py trunc_normal_: bounds are the ABSOLUTE values a=-2, b=2 (not +-2 std)."""
    return nn.init.trunc_normal_(t, std=std, a=-2., b=2.)


class VisionTransformer(nn.Module):
    """timm.models.VisionTransformer (vision_transformer.py:711-1302), the subset the kernels cover:
    learned absolute pos-embed, class / register tokens, token or avg pooling, LayerNorm, exact GELU."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, global_pool="token",
                 embed_dim=768, depth=12, num_heads=12, mlp_ratio=4., qkv_bias=True, qk_norm=False,
                 scale_attn_norm=False, scale_mlp_norm=False, proj_bias=True, init_values=None, class_token=True,
                 pos_embed="learn", no_embed_class=False, reg_tokens=0, pre_norm=False, final_norm=True,
                 fc_norm=None, pool_include_prefix=False, dynamic_img_size=False, dynamic_img_pad=False,
                 drop_rate=0., pos_drop_rate=0., patch_drop_rate=0., proj_drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0., weight_init="", fix_init=False, embed_layer=PatchEmbed, embed_norm_layer=None,
                 norm_layer=None, act_layer=None, block_fn=Block, mlp_layer=Mlp, attn_layer=Attention,
                 device=None, dtype=None):
        super().__init__()
        dd = {"device": device, "dtype": dtype}
        assert global_pool in ("", "avg", "token", "map"), f"global_pool={global_pool!r} is not supported by b200timm"
        assert class_token or global_pool != "token"
        assert pos_embed in ("", "none", "learn")
        if dynamic_img_pad or patch_drop_rate > 0 or embed_norm_layer is not None:
            raise B200Error("dynamic_img_pad / patch dropout / embed norm are not supported")
        use_fc_norm = global_pool == "avg" if fc_norm is None else fc_norm
        if isinstance(norm_layer, str):       # create_norm._NORM_MAP keys (timm/layers/create_norm.py:31-50)
            from .layers import RmsNorm
            known = {"layernorm": LayerNorm, "rmsnorm": RmsNorm}
            if norm_layer not in known:
                raise B200Error(f"norm_layer {norm_layer!r} is not covered (layernorm, rmsnorm)")
            norm_layer = known[norm_layer]
        if isinstance(act_layer, str):        # create_act keys (timm/layers/create_act.py); 'silu' goes with the gated MLPs
            known = {"gelu": nn.GELU, "gelu_tanh": partial(nn.GELU, approximate="tanh"), "silu": nn.SiLU, "swish": nn.SiLU,
                     "sigmoid": nn.Sigmoid}
            if act_layer not in known:
                raise B200Error(f"act_layer {act_layer!r} is not covered (gelu, gelu_tanh, silu, sigmoid)")
            act_layer = known[act_layer]
        if isinstance(mlp_layer, str):        # the reference has no string lookup for mlp_layer (vision_transformer.py:758);
            from . import layers as _layers   # names of timm.layers classes are accepted here so fixtures / CLIs can say it
            known = {"Mlp": Mlp, "SwiGLU": _layers.SwiGLU, "SwiGLUPacked": _layers.SwiGLUPacked, "GluMlp": _layers.GluMlp}
            if mlp_layer not in known:
                raise B200Error(f"mlp_layer {mlp_layer!r} is not covered ({', '.join(known)})")
            mlp_layer = known[mlp_layer]
        norm_layer = norm_layer or LayerNorm
        act_layer = act_layer or nn.GELU
        self.num_classes = num_classes
        self.in_chans = in_chans
        self.global_pool = global_pool
        self.num_features = self.head_hidden_size = self.embed_dim = embed_dim
        self.num_prefix_tokens = (1 if class_token else 0) + reg_tokens
        self.num_reg_tokens = reg_tokens
        self.has_class_token = class_token
        self.no_embed_class = no_embed_class
        self.pool_include_prefix = pool_include_prefix
        self.dynamic_img_size = dynamic_img_size
        self.grad_checkpointing = False
        self.attn_pool = None       # replaced below for global_pool='map' (vision_transformer.py:882-890)

        embed_args = dict(strict_img_size=False) if dynamic_img_size else {}     # vision_transformer.py:818-821
        self.patch_embed = embed_layer(img_size=img_size, patch_size=patch_size, in_chans=in_chans,
                                       embed_dim=embed_dim, bias=not pre_norm, **embed_args, **dd)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.empty(1, 1, embed_dim, **dd)) if class_token else None
        self.reg_token = nn.Parameter(torch.empty(1, reg_tokens, embed_dim, **dd)) if reg_tokens else None
        embed_len = num_patches if no_embed_class else num_patches + self.num_prefix_tokens
        if not pos_embed or pos_embed == "none":
            self.pos_embed = None
        else:
            self.pos_embed = nn.Parameter(torch.empty(1, embed_len, embed_dim, **dd))
        self.pos_drop = nn.Dropout(p=pos_drop_rate)
        self.patch_drop = nn.Identity()
        self.norm_pre = norm_layer(embed_dim, **dd) if pre_norm else nn.Identity()
        dpr = calculate_drop_path_rates(drop_path_rate, depth)
        self.blocks = nn.Sequential(*[
            block_fn(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_norm=qk_norm,
                     scale_attn_norm=scale_attn_norm, scale_mlp_norm=scale_mlp_norm, proj_bias=proj_bias,
                     init_values=init_values, proj_drop=proj_drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i],
                     norm_layer=norm_layer, act_layer=act_layer, mlp_layer=mlp_layer, attn_layer=attn_layer,
                     depth=i, **dd)
            for i in range(depth)])
        self.norm = norm_layer(embed_dim, **dd) if final_norm and not use_fc_norm else nn.Identity()
        if global_pool == "map":
            from .layers import AttentionPoolLatent
            self.attn_pool = AttentionPoolLatent(embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, norm_layer=norm_layer,
                                                 act_layer=act_layer, **dd)
        self.fc_norm = norm_layer(embed_dim, **dd) if final_norm and use_fc_norm else nn.Identity()
        self.head_drop = nn.Dropout(drop_rate)
        self.head = nn.Linear(embed_dim, num_classes, **dd) if num_classes > 0 else nn.Identity()
        if weight_init != "skip":
            self.init_weights()

    def init_weights(self, mode: str = "", needs_reset: bool = False):
        """Default timm ViT init (vision_transformer.py:920-938, init_weights_vit_timm :1305-1320)."""
        if self.pos_embed is not None:
            trunc_normal_(self.pos_embed, std=.02)
        if self.cls_token is not None:
            nn.init.normal_(self.cls_token, std=1e-6)
        if self.reg_token is not None:
            nn.init.normal_(self.reg_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    @torch.jit.ignore
    def set_grad_checkpointing(self, enable: bool = True):
        """timm's switch (vision_transformer.py:983-990): recompute each block in the backward instead of keeping its
        activations (functional.vit_forward_features)."""
        self.grad_checkpointing = enable

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"pos_embed", "cls_token", "dist_token"}

    def get_classifier(self):
        return self.head

    def forward_features(self, x, attn_mask=None, is_causal=False, kv_len=None):
        if kv_len is None and (attn_mask is not None or is_causal):
            from .layers import attn_mask_operand
            kv_len = attn_mask_operand(attn_mask, is_causal, self.patch_embed.num_patches + self.num_prefix_tokens, x.shape[0],
                                       x.device)
        return BF.vit_forward_features(self, x, kv_len)

    def forward_head(self, x, pre_logits: bool = False):
        return BF.vit_forward_head(self, x, pre_logits)

    def forward(self, x, attn_mask=None, is_causal=False, kv_len=None):
        if kv_len is None and (attn_mask is not None or is_causal):
            from .layers import attn_mask_operand
            kv_len = attn_mask_operand(attn_mask, is_causal, self.patch_embed.num_patches + self.num_prefix_tokens, x.shape[0],
                                       x.device)
        return BF.vit_forward(self, x, kv_len)


# name -> constructor kwargs, copied from the reference entrypoints' model_args
MODEL_ARGS = {
    "vit_tiny_patch16_224": dict(patch_size=16, embed_dim=192, depth=12, num_heads=3),            # :3105-3110
    "vit_small_patch16_224": dict(patch_size=16, embed_dim=384, depth=12, num_heads=6),
    "vit_base_patch16_224": dict(patch_size=16, embed_dim=768, depth=12, num_heads=12),           # :3188
    "vit_large_patch16_224": dict(patch_size=16, embed_dim=1024, depth=24, num_heads=16),
    "vit_large_patch14_clip_336": dict(img_size=336, patch_size=14, embed_dim=1024, depth=24, num_heads=16,
                                       pre_norm=True, norm_layer=partial(LayerNorm, eps=1e-5)),  # :3524-3531
}


def create_model(name: str, **kwargs) -> VisionTransformer:
    """Resolve one of the benchmark model names (stand-in for timm.create_model, _factory.py:44-149)."""
    if name not in MODEL_ARGS:
        raise B200Error(f"unknown model {name!r}; b200timm mirrors {sorted(MODEL_ARGS)} -- for any other timm model "
                        "use timm.create_model together with b200timm.plugin")
    return VisionTransformer(**dict(MODEL_ARGS[name], **kwargs))
