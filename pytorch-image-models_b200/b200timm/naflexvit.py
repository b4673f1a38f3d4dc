"""Mirror of timm's NaFlexVit for BASELINE.json configs[4] (vit_large_patch14_clip_336 run in NaFlex mode):
pre-patchified variable-length input, learned 2-D position table interpolated per sample, key-padding attention.

Reference: timm/models/naflexvit.py -- NaFlexEmbeds :321-960 (linear patch projection :449,905-908, per-sample
learned position embedding :567-621, prefix tokens :944-953), create_attention_mask :964-1029,
NaFlexVit.forward_features :1711-1773, forward_head / global_pool_naflex :1033-1102,1775-1814, forward :1816-1864.
Module / parameter names follow timm (``embeds.{cls_token,reg_token,pos_embed,proj}``, ``norm_pre``, ``blocks.*``,
``norm``, ``head``) so NaFlexVit checkpoints load with strict=True.

B200 execution: patch rows [B*N, P*P*C] -> (zero-padded to a 16-byte row pitch) -> GEMM -> naflex_assemble kernel
(adds the interpolated position rows, prepends prefix tokens) -> LayerNorm -> transformer blocks whose attention
takes the key-padding mask as per-sample key counts (the [B,1,N,N] additive mask of :964-1029 is never built) ->
final LayerNorm on the pooled token -> head GEMM.  The bicubic interpolation of the small [gh, gw, D] position table
(one per distinct grid size, :592-604) stays in PyTorch, as in the reference.
"""
from typing import Dict, Optional, Tuple, Union

import torch
import torch.nn.functional as F
from torch import nn

from . import functional as BF
from . import ops
from ._lib import B200Error
from .layers import LayerNorm
from .vision_transformer import Block, calculate_drop_path_rates, trunc_normal_


class NaFlexEmbeds(nn.Module):
    """Parameter holder mirroring timm's NaFlexEmbeds (linear projection, learned position table)."""

    def __init__(self, patch_size=16, in_chans=3, embed_dim=768, proj_bias=True, class_token=True, reg_tokens=0,
                 pos_embed_grid_size=(16, 16), device=None, dtype=None):
        super().__init__()
        dd = {"device": device, "dtype": dtype}
        self.patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.embed_dim = embed_dim
        self.num_prefix_tokens = (1 if class_token else 0) + reg_tokens
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim, **dd)) if class_token else None
        self.reg_token = nn.Parameter(torch.zeros(1, reg_tokens, embed_dim, **dd)) if reg_tokens else None
        self.pos_embed = nn.Parameter(torch.empty(1, pos_embed_grid_size[0], pos_embed_grid_size[1], embed_dim, **dd))
        self.proj = nn.Linear(self.patch_size[0] * self.patch_size[1] * in_chans, embed_dim, bias=proj_bias, **dd)
        trunc_normal_(self.pos_embed, std=.02)
        if self.cls_token is not None:
            nn.init.normal_(self.cls_token, std=1e-6)

    def position_rows(self, sizes, dtype, device):
        """Concatenated position rows for the distinct (h, w) grids in `sizes` (naflexvit.py:590-604) and, per
        sample, the row offset / row count into that table."""
        gh, gw = self.pos_embed.shape[1:3]
        nchw = None
        tables, offsets, off = {}, [], 0
        chunks = []
        for s in sizes:
            if s not in tables:
                if s == (gh, gw):
                    t = self.pos_embed.reshape(gh * gw, -1)
                else:
                    if nchw is None:
                        nchw = self.pos_embed.permute(0, 3, 1, 2).float()
                    t = F.interpolate(nchw, size=s, mode="bicubic", align_corners=False, antialias=True)
                    t = t.flatten(2).transpose(1, 2)[0]
                tables[s] = off
                chunks.append(t.to(dtype))
                off += s[0] * s[1]
            offsets.append(tables[s])
        table = torch.cat(chunks, 0).contiguous()
        pos_off = torch.tensor(offsets, dtype=torch.int32, device=device)
        pos_len = torch.tensor([h * w for h, w in sizes], dtype=torch.int32, device=device)
        return table, pos_off, pos_len


class NaFlexVit(nn.Module):
    """timm.models.naflexvit.NaFlexVit for learned position embeddings, standard attention, token / avg pooling."""

    def __init__(self, patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0, qkv_bias=True,
                 class_token=True, reg_tokens=0, pos_embed_grid_size=(16, 16), pre_norm=False, final_norm=True,
                 global_pool="token", init_values=None, drop_path_rate=0., norm_eps=1e-6, in_chans=3, num_classes=1000,
                 device=None, dtype=None):
        super().__init__()
        dd = {"device": device, "dtype": dtype}
        if global_pool not in ("token", "avg"):
            raise B200Error(f"NaFlexVit global_pool={global_pool!r} is not covered (token / avg)")
        self.num_classes, self.global_pool, self.embed_dim = num_classes, global_pool, embed_dim
        self.num_prefix_tokens = (1 if class_token else 0) + reg_tokens
        self.in_chans = in_chans
        norm_layer = lambda d, **kw: LayerNorm(d, eps=norm_eps, **kw)   # noqa: E731
        self.embeds = NaFlexEmbeds(patch_size, in_chans, embed_dim, proj_bias=not pre_norm, class_token=class_token,
                                   reg_tokens=reg_tokens, pos_embed_grid_size=pos_embed_grid_size, **dd)
        self.norm_pre = norm_layer(embed_dim, **dd) if pre_norm else nn.Identity()
        dpr = calculate_drop_path_rates(drop_path_rate, depth)
        self.blocks = nn.Sequential(*[
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, init_values=init_values,
                  drop_path=dpr[i], norm_layer=norm_layer, depth=i, **dd) for i in range(depth)])
        use_fc_norm = global_pool == "avg"
        self.norm = norm_layer(embed_dim, **dd) if final_norm and not use_fc_norm else nn.Identity()
        self.fc_norm = norm_layer(embed_dim, **dd) if final_norm and use_fc_norm else nn.Identity()
        self.head_drop = nn.Dropout(0.)
        self.head = nn.Linear(embed_dim, num_classes, **dd) if num_classes > 0 else nn.Identity()
        for m in self.modules():
            if isinstance(m, nn.Linear):
                trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    # -- embeds + blocks --------------------------------------------------------------------------------------
    def _tokens(self, patches, patch_coord, patch_valid, rows=None):
        """rows: the projection's operand already built (bf16 [B * N, Kpad], _image_to_patches); else built from `patches`."""
        if not patches.is_cuda:
            raise B200Error("libb200timm kernels need CUDA tensors (there is no CPU path)")
        e = self.embeds
        B, N = patches.shape[:2]
        train = torch.is_grad_enabled() and any(p.requires_grad for p in e.parameters())
        w = BF._w(e.proj)
        K = w.shape[1]
        kpad = (K + 7) // 8 * 8
        if rows is not None:
            x = rows
        elif patches.dtype == torch.uint8:
            # the loader's raw uint8 patches (b200timm.data.NaFlexPrefetchLoader): normalisation, bf16 cast and the K padding
            # happen in ONE pass (b200_patches_u8) instead of the fp32 sub / div round trip of naflex_loader.py:161
            x = ops.patches_u8(patches.contiguous(), self.in_chans, channels_last=True, kpad=kpad)
        else:
            x = patches.reshape(B * N, -1)
            if x.dtype == torch.float32 and e.proj.weight.dtype == torch.float32:
                x = x.to(torch.bfloat16)        # train.py --amp: the cast autocast applies to F.linear's input
            if x.dtype != torch.bfloat16:
                raise B200Error("NaFlexVit B200 path needs bf16 patches (fp32 patches only with fp32 master weights) or the "
                                "loader's uint8 patches")
            if kpad != K:   # p=14: 588 -> 592, 16-byte row pitch for TMA (zero columns contribute nothing)
                x = F.pad(x, (0, kpad - K))
        if kpad != K:
            if not train:
                cache = getattr(e, "_b200_wpad", None)
                if cache is None or cache[0] != ops.wkey(w):
                    wp = torch.zeros(w.shape[0], kpad, dtype=w.dtype, device=w.device)
                    wp[:, :K] = w.detach()
                    e._b200_wpad = (ops.wkey(w), wp)
                    cache = e._b200_wpad
                w = cache[1]
        x = x.contiguous()
        # per-sample grid size from the coordinates (calculate_naflex_grid_sizes, naflexvit.py:204-208)
        sizes = [(int(h), int(w_)) for h, w_ in (patch_coord.amax(dim=1) + 1).tolist()]
        table, pos_off, pos_len = e.position_rows(sizes, torch.bfloat16, patches.device)
        pre = [t for t in (e.cls_token, e.reg_token) if t is not None]
        prefix = None
        if pre:
            prefix = (pre[0] if len(pre) == 1 else torch.cat(pre, 1)).reshape(self.num_prefix_tokens, -1).contiguous()
        if train:
            from .autograd import NaflexEmbedFn
            tokens = NaflexEmbedFn.apply(x, w, e.proj.bias, prefix, table, pos_off, pos_len, B, N)
        else:
            proj = ops.gemm(x, w, B * N, w.shape[0], kpad, bias=e.proj.bias)
            tokens = ops.naflex_assemble(proj, prefix, table, pos_off, pos_len, B, N, self.num_prefix_tokens,
                                         self.embed_dim)
        # key-padding: prefix tokens + the valid (prefix-contiguous) patches (create_attention_mask, :964-1029)
        nvalid = patch_valid.sum(1).to(torch.int32)
        expect = torch.arange(N, device=patches.device)[None, :] < nvalid[:, None]
        if not bool((patch_valid == expect).all()):
            raise B200Error("patch_valid must be prefix-contiguous (NaFlex collator layout, naflex_dataset.py:136-142)")
        kv_len = (nvalid + self.num_prefix_tokens).contiguous()
        return tokens, kv_len

    def _image_to_patches(self, img):
        """An NCHW image batch as NaFlexEmbeds.forward takes it (naflexvit.py:880-890): batch_patchify fused into the gather
        that builds the projection's GEMM operand (b200_batch_patchify: P-P-C patch layout, ragged edges zero-padded, uint8
        pixels normalised on the fly) + the regular grid of coordinates; every patch is valid."""
        P = self.embeds.patch_size
        P = P[0] if isinstance(P, (tuple, list)) else P
        B = img.shape[0]
        K = img.shape[1] * P * P
        rows, (gh, gw) = ops.batch_patchify(img.contiguous(), P, channels_last=True, kpad=(K + 7) // 8 * 8)
        ys, xs = torch.meshgrid(torch.arange(gh, device=img.device), torch.arange(gw, device=img.device), indexing="ij")
        coord = torch.stack([ys, xs], -1).reshape(1, gh * gw, 2).expand(B, -1, -1).contiguous()
        valid = torch.ones(B, gh * gw, dtype=torch.bool, device=img.device)
        return rows, coord, valid

    def forward_features(self, patches, patch_coord=None, patch_valid=None) -> Dict[str, torch.Tensor]:
        rows = None
        if patches.dim() == 4 and patch_coord is None:
            rows, patch_coord, patch_valid = self._image_to_patches(patches)
            patches = rows.view(patches.shape[0], -1, rows.shape[1])
        if patch_coord is None or patch_valid is None:
            raise B200Error("NaFlexVit mirror needs (patches, patch_coord, patch_valid) or an NCHW image batch")
        x, kv_len = self._tokens(patches, patch_coord, patch_valid, rows)
        if not isinstance(self.norm_pre, nn.Identity):
            x = BF.layernorm(self.norm_pre, x)
        carry = BF.LnCarry()
        for blk in self.blocks:
            x = BF.block_forward(blk, x, kv_len, carry)
        if not isinstance(self.norm, nn.Identity):
            x = BF.layernorm(self.norm, x)
        return {"patches": x, "patch_valid": patch_valid}

    def forward_head(self, x, patch_valid=None, pre_logits=False):
        if isinstance(x, dict):
            x, patch_valid = x["patches"], x["patch_valid"]
        if self.global_pool == "token":
            x = x[:, 0]
        else:   # masked average over the valid patch tokens (global_pool_naflex, :1077-1082)
            m = patch_valid.unsqueeze(-1).float()
            x = ((x[:, self.num_prefix_tokens:].float() * m).sum(1) / m.sum(1).clamp(min=1)).to(x.dtype)
        if not isinstance(self.fc_norm, nn.Identity):
            x = BF.layernorm(self.fc_norm, x.contiguous())
        if pre_logits or isinstance(self.head, nn.Identity):
            return x
        x = x.contiguous()
        if BF._needs_grad(x, self.head):
            from .autograd import LinearFn
            return LinearFn.apply(x, BF._w(self.head), self.head.bias)
        if self.head.weight.shape[0] % 8:
            BF._w(self.head)
            return BF._padded_out_linear(self.head, x)
        return ops.linear_fwd(x, BF._w(self.head), self.head.bias)

    def forward(self, patches: Union[torch.Tensor, Dict[str, torch.Tensor]], patch_coord=None, patch_valid=None):
        if isinstance(patches, dict):
            patches, patch_coord, patch_valid = patches["patches"], patches["patch_coord"], patches["patch_valid"]
        return self.forward_head(self.forward_features(patches, patch_coord, patch_valid))


MODEL_ARGS = {
    # vit_large_patch14_clip_336 (vision_transformer.py:3524-3531) through the classic->NaFlex shim (:2034-2074):
    # 24x24 position grid, class token, pre-norm, LayerNorm eps 1e-5, token pooling
    "vit_large_patch14_clip_336": dict(patch_size=14, embed_dim=1024, depth=24, num_heads=16, class_token=True,
                                       pos_embed_grid_size=(24, 24), pre_norm=True, norm_eps=1e-5, global_pool="token"),
}


def create_model(name: str, **kwargs) -> NaFlexVit:
    if name not in MODEL_ARGS:
        raise B200Error(f"unknown NaFlex variant {name!r}; mirrored: {sorted(MODEL_ARGS)}")
    return NaFlexVit(**dict(MODEL_ARGS[name], **kwargs))
