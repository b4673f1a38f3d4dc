"""ORACLE (test infrastructure, NOT product code): CPU fp32 restatement of timm's NaFlexVit forward in NaFlex mode
(pre-patchified input, learned position table, key-padding mask, token pooling).  Pinned against
tests/golden/naflex_g1.pt (made from the unmodified reference by oracle/make_golden.py)."""
import torch
import torch.nn.functional as F

from .vit_oracle import block, layer_norm, linear


def naflex_embeds(patches, patch_coord, sd, n_prefix_keys=("embeds.cls_token", "embeds.reg_token")):
    """NaFlexEmbeds.forward, timm/models/naflexvit.py:857-960 (linear mode, learned position embedding)."""
    x = linear(patches.flatten(2), sd["embeds.proj.weight"], sd.get("embeds.proj.bias"))            # :905-908
    pos = sd["embeds.pos_embed"]                                                                     # [1, gh, gw, D]
    gh, gw = pos.shape[1:3]
    x = x.clone()
    for b in range(x.shape[0]):                                                                      # :567-621
        h = int(patch_coord[b, :, 0].max()) + 1
        w = int(patch_coord[b, :, 1].max()) + 1
        if (h, w) == (gh, gw):
            flat = pos.reshape(gh * gw, -1)
        else:
            flat = F.interpolate(pos.permute(0, 3, 1, 2).float(), size=(h, w), mode="bicubic", align_corners=False,
                                 antialias=True).flatten(2).transpose(1, 2)[0]
        n = min(x.shape[1], h * w)
        x[b, :n] += flat[:n]
    to_cat = [sd[k].expand(x.shape[0], -1, -1) for k in n_prefix_keys if k in sd]                    # :944-953
    return torch.cat(to_cat + [x], 1), sum(sd[k].shape[1] for k in n_prefix_keys if k in sd)


def naflex_forward(patches, patch_coord, patch_valid, sd, cfg):
    """NaFlexVit.forward, naflexvit.py:1711-1864 (token pooling)."""
    eps = cfg.get("eps", 1e-6)
    x, npre = naflex_embeds(patches, patch_coord, sd)
    kv_len = patch_valid.sum(1) + npre                      # create_attention_mask for prefix-valid padding, :964-1029
    if "norm_pre.weight" in sd:
        x = layer_norm(x, sd["norm_pre.weight"], sd["norm_pre.bias"], eps)
    for i in range(cfg["depth"]):
        x = block(x, sd, f"blocks.{i}.", cfg["num_heads"], eps, kv_len)
    if "norm.weight" in sd:
        x = layer_norm(x, sd["norm.weight"], sd["norm.bias"], eps)
    feats = x
    x = x[:, 0]
    return feats, linear(x, sd["head.weight"], sd.get("head.bias"))
