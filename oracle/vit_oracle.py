"""ORACLE (test infrastructure, NOT product code): CPU fp32 restatement of timm's ViT hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker / CPU baseline -- never as the thing shipped.  The product path lives in
pytorch-image-models_b200/ and fails loudly when its CUDA library is missing.

Parity pinning: every function below restates one reference function (file:line relative to the timm
checkout, timm 1.0.28.dev0) and is pinned against golden vectors produced by the *unmodified* reference
in the authoring container (oracle/make_golden.py -> tests/golden/*.pt, checked by
tests/test_oracle_golden.py).  The reference's own golden test (tests/test_models.py:132-173) needs
network access for pretrained weights and cannot run here, so those fixtures are the pin.

The arithmetic is written out explicitly with torch CPU fp32 tensor ops (matmul / exp / mean), i.e. it does
not call F.layer_norm, F.scaled_dot_product_attention or F.gelu, so that it is an independent statement of
what the CUDA kernels must compute.
"""
import math

import torch


# ----------------------------------------------------------------------------------------------- leaf ops
def layer_norm(x, weight, bias, eps):
    """timm/layers/norm.py:85-90 (LayerNorm.forward -> F.layer_norm): biased variance over the last dim."""
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    y = (x - mean) / torch.sqrt(var + eps)
    if weight is not None:
        y = y * weight
    if bias is not None:
        y = y + bias
    return y


def rms_norm(x, weight, eps):
    """timm/layers/fast_norm.py:119-139 (rms_norm; RmsNorm.forward norm.py:241-248): x * rsqrt(mean(x^2) + eps) * w."""
    v = (x * x).mean(-1, keepdim=True)
    y = x * torch.rsqrt(v + eps)
    return y if weight is None else y * weight


def norm(x, sd, prefix, eps):
    """LayerNorm, or RMSNorm when the state dict holds no bias for this norm (timm's RmsNorm has only `weight`)."""
    if prefix + "bias" in sd:
        return layer_norm(x, sd[prefix + "weight"], sd[prefix + "bias"], eps)
    return rms_norm(x, sd[prefix + "weight"], eps)


def gelu_erf(x):
    """nn.GELU() default (exact erf form), timm/layers/mlp.py:24,49."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def linear(x, weight, bias=None):
    """nn.Linear: y = x W^T + b (timm/layers/attention.py:100,105; mlp.py:40,44)."""
    y = x @ weight.t()
    return y if bias is None else y + bias


def attention(x, qkv_w, qkv_b, proj_w, proj_b, num_heads, kv_len=None, norms=None):
    """timm/layers/attention.py:108-139 (Attention.forward, no dropout).  norms = (sd, prefix, eps): the optional
    q_norm / k_norm over head_dim (:117) and norm over the merged heads (:136) are applied when their weights are in sd
    (LayerNorm, or RMSNorm when no bias key exists).

    kv_len: optional int tensor [B]; keys >= kv_len[b] are masked out, which is what the additive
    finfo.min mask built by naflexvit.py:964-1029 does for prefix-valid padding.
    """
    B, N, C = x.shape
    hd = qkv_w.shape[0] // 3 // num_heads
    qkv = linear(x, qkv_w, qkv_b).reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)  # :115
    q, k, v = qkv.unbind(0)                                                                 # :116
    if norms is not None and norms[1] + "q_norm.weight" in norms[0]:                        # :117
        q = norm(q, norms[0], norms[1] + "q_norm.", norms[2])
        k = norm(k, norms[0], norms[1] + "k_norm.", norms[2])
    s = (q * hd ** -0.5) @ k.transpose(-2, -1)                                             # :127-128
    if kv_len is not None:
        keys = torch.arange(N)[None, None, None, :]
        s = s.masked_fill(keys >= kv_len.view(B, 1, 1, 1), float("-inf"))
    s = s - s.amax(-1, keepdim=True)
    p = torch.exp(s)
    p = p / p.sum(-1, keepdim=True)                                                         # :131 softmax
    o = (p @ v).transpose(1, 2).reshape(B, N, num_heads * hd)                               # :133,135
    if norms is not None and norms[1] + "norm.weight" in norms[0]:                          # :136 (scale_norm)
        o = norm(o, norms[0], norms[1] + "norm.", norms[2])
    return linear(o, proj_w, proj_b)                                                        # :137


def rope_rotate(x, emb, half=False):
    """apply_rot_embed_cat (timm/layers/pos_embed_sincos.py:281-297): emb = [sin | cos] along the last dim,
    x * cos + rot(x) * sin with rot = interleaved pairs (`rot`, :227-230: [-x1, x0, -x3, x2, ...]) or, for half=True,
    `rope_rotate_half` (:233-237: [-x_{d/2..}, x_{..d/2}])."""
    d = x.shape[-1]
    sin, cos = emb[..., :d], emb[..., d:]
    if half:
        r = torch.cat([-x[..., d // 2:], x[..., :d // 2]], dim=-1)
    else:
        r = torch.empty_like(x)
        r[..., 0::2] = -x[..., 1::2]
        r[..., 1::2] = x[..., 0::2]
    return x * cos + r * sin


def attention_rope(x, qkv_w, qkv_b, proj_w, proj_b, num_heads, rope, num_prefix=1, half=False, kv_len=None,
                   round_qkv=None):
    """timm/layers/attention.py:224-282 (AttentionRope.forward, fused qkv, no qk-norm / scale-norm / dropout):
    q, k of every token after the first `num_prefix` are rotated by `rope` ([N - num_prefix, 2 * head_dim]) before the
    scaled-dot-product attention (:249-253).  round_qkv: optional dtype the Linear output is rounded to before the
    rotation and again after it (what a bf16 module does: nn.Linear emits bf16, `.type_as(v)` rounds the result)."""
    B, N, C = x.shape
    hd = qkv_w.shape[0] // 3 // num_heads
    qkv = linear(x, qkv_w, qkv_b)
    if round_qkv is not None:
        qkv = qkv.to(round_qkv).to(x.dtype)
    qkv = qkv.reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)                        # :236-237
    q, k, v = qkv.unbind(0)
    if rope is not None:
        def rot(t):
            out = torch.cat([t[:, :, :num_prefix], rope_rotate(t[:, :, num_prefix:], rope.to(t.dtype), half)], dim=2)
            return out if round_qkv is None else out.to(round_qkv).to(t.dtype)
        q, k = rot(q), rot(k)                                                               # :252-253
    s = (q * hd ** -0.5) @ k.transpose(-2, -1)
    if kv_len is not None:
        keys = torch.arange(N)[None, None, None, :]
        s = s.masked_fill(keys >= kv_len.view(B, 1, 1, 1), float("-inf"))
    s = s - s.amax(-1, keepdim=True)
    p = torch.exp(s)
    p = p / p.sum(-1, keepdim=True)
    o = (p @ v).transpose(1, 2).reshape(B, N, num_heads * hd)
    return linear(o, proj_w, proj_b)


def gelu_tanh(x):
    """timm GELUTanh / nn.GELU(approximate='tanh') (timm/layers/activations.py): 0.5 x (1 + tanh(sqrt(2/pi)(x + 0.044715 x^3)))."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


ACTS = {"gelu": gelu_erf, "gelu_tanh": gelu_tanh, "silu": lambda t: t * torch.sigmoid(t)}


def mlp(x, fc1_w, fc1_b, fc2_w, fc2_b, act="gelu"):
    """timm/layers/mlp.py:47-54 (Mlp.forward, drop=0, norm=Identity); act = the create_act.py key of act_layer."""
    return linear(ACTS[act](linear(x, fc1_w, fc1_b)), fc2_w, fc2_b)


def glu_mlp(x, sd, prefix, eps=1e-6, act="silu", gate_last=False):
    """Gated MLPs, timm/layers/mlp.py.  SwiGLU.forward (:153-162, keys fc1_g / fc1_x): fc2(norm(act(fc1_g x) * fc1_x x)).
    GluMlp.forward (:104-113, key fc1): x1, x2 = fc1(x).chunk(2); x1 * act(x2) if gate_last else act(x1) * x2
    (SwiGLUPacked :116 = SiLU, gate_last=False).  The inner norm is applied when its weights are in `sd`."""
    f = {"silu": lambda t: t * torch.sigmoid(t), "sigmoid": torch.sigmoid}[act]
    g = lambda k: sd.get(prefix + k)  # noqa: E731
    if g("fc1_g.weight") is not None:
        h = f(linear(x, g("fc1_g.weight"), g("fc1_g.bias"))) * linear(x, g("fc1_x.weight"), g("fc1_x.bias"))
    else:
        x1, x2 = linear(x, g("fc1.weight"), g("fc1.bias")).chunk(2, dim=-1)
        h = x1 * f(x2) if gate_last else f(x1) * x2
    if g("norm.weight") is not None:
        h = layer_norm(h, g("norm.weight"), g("norm.bias"), eps)
    return linear(h, g("fc2.weight"), g("fc2.bias"))


def attention_pool_latent(x, sd, prefix, num_heads, eps=1e-6, act="gelu"):
    """AttentionPoolLatent.forward, timm/layers/attention_pool.py:95-128 (latent_len 1, no qk_norm / pos_embed, token pool):
    q = q(latent), k, v = kv(x); softmax(q k^T / sqrt(hd)) v; proj; x + mlp(norm(x)); x[:, 0]."""
    B, N, C = x.shape
    hd = C // num_heads
    g = lambda k: sd.get(prefix + k)  # noqa: E731
    q = linear(g("latent").expand(B, -1, -1), g("q.weight"), g("q.bias")).reshape(B, 1, num_heads, hd).transpose(1, 2)
    kv = linear(x, g("kv.weight"), g("kv.bias")).reshape(B, N, 2, num_heads, hd).permute(2, 0, 3, 1, 4)
    k, v = kv.unbind(0)
    s = (q * hd ** -0.5) @ k.transpose(-2, -1)
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, 1, C)
    o = linear(o, g("proj.weight"), g("proj.bias"))
    h = norm(o, sd, prefix + "norm.", eps) if g("norm.weight") is not None else o
    o = o + mlp(h, g("mlp.fc1.weight"), g("mlp.fc1.bias"), g("mlp.fc2.weight"), g("mlp.fc2.bias"), act)
    return o[:, 0]


def drop_path(x, keep_mask, keep_prob, scale_by_keep=True):
    """timm/layers/drop.py:158-175 with the Bernoulli draw made explicit: keep_mask is a 0/1 tensor [B]."""
    m = keep_mask.to(x.dtype).view(-1, *([1] * (x.dim() - 1)))
    if keep_prob > 0.0 and scale_by_keep:
        m = m / keep_prob
    return x * m


def patch_embed(img, weight, bias, patch):
    """timm/layers/patch_embed.py:117-142: Conv2d(k=stride=patch) -> flatten(2).transpose(1,2).

    Written as an explicit patch gather + matmul (k = c*p*p + ky*p + kx, Conv2d.weight.flatten(1) order).
    """
    B, C, H, W = img.shape
    gh, gw = H // patch, W // patch
    x = img[:, :, :gh * patch, :gw * patch].reshape(B, C, gh, patch, gw, patch)
    x = x.permute(0, 2, 4, 1, 3, 5).reshape(B, gh * gw, C * patch * patch)
    return linear(x, weight.reshape(weight.shape[0], -1), bias)


def resample_abs_pos_embed(pos, new_size, old_size, num_prefix_tokens):
    """timm/layers/pos_embed.py:19-59: unchanged when the grid already matches (square); otherwise the grid rows go
    through F.interpolate(bicubic, antialias=True) in fp32 as [1, D, gh, gw] and the prefix rows are put back."""
    import torch.nn.functional as F
    if new_size[0] * new_size[1] + num_prefix_tokens == pos.shape[1] and new_size[0] == new_size[1]:
        return pos
    prefix, grid = pos[:, :num_prefix_tokens], pos[:, num_prefix_tokens:]
    D = pos.shape[-1]
    g = F.interpolate(grid.float().reshape(1, old_size[0], old_size[1], D).permute(0, 3, 1, 2), size=tuple(new_size),
                      mode="bicubic", antialias=True)
    g = g.permute(0, 2, 3, 1).reshape(1, -1, D).to(pos.dtype)
    return torch.cat([prefix, g], 1) if num_prefix_tokens else g


# ----------------------------------------------------------------------------------------------- model
def block(x, sd, prefix, num_heads, eps, kv_len=None, keep1=None, keep2=None, keep_prob=1.0, glu=None, act="gelu"):
    """timm/models/vision_transformer.py:206-214 (Block.forward).  glu: None (Mlp) or dict(act=, gate_last=) for a
    gated mlp_layer (SwiGLU is recognised by its fc1_g key)."""
    g = lambda k: sd.get(prefix + k)  # noqa: E731
    h = attention(norm(x, sd, prefix + "norm1.", eps), g("attn.qkv.weight"), g("attn.qkv.bias"),
                  g("attn.proj.weight"), g("attn.proj.bias"), num_heads, kv_len, norms=(sd, prefix + "attn.", eps))
    if g("ls1.gamma") is not None:
        h = h * g("ls1.gamma")                                   # timm/layers/layer_scale.py:26-27
    if keep1 is not None:
        h = drop_path(h, keep1, keep_prob)
    x = x + h
    if glu is not None or g("mlp.fc1_g.weight") is not None:
        h = glu_mlp(norm(x, sd, prefix + "norm2.", eps), sd, prefix + "mlp.", eps, **(glu or {}))
    else:
        h = mlp(norm(x, sd, prefix + "norm2.", eps), g("mlp.fc1.weight"), g("mlp.fc1.bias"),
                g("mlp.fc2.weight"), g("mlp.fc2.bias"), act)
    if g("ls2.gamma") is not None:
        h = h * g("ls2.gamma")
    if keep2 is not None:
        h = drop_path(h, keep2, keep_prob)
    return x + h


def vit_forward_features(img, sd, cfg, kv_len=None, return_blocks=False):
    """timm/models/vision_transformer.py:1231-1253 (forward_features) incl. _pos_embed :1038-1075."""
    eps = cfg.get("eps", 1e-6)
    x = patch_embed(img, sd["patch_embed.proj.weight"], sd.get("patch_embed.proj.bias"), cfg["patch_size"])
    B = x.shape[0]
    to_cat = []
    if "cls_token" in sd:
        to_cat.append(sd["cls_token"].expand(B, -1, -1))
    if "reg_token" in sd:
        to_cat.append(sd["reg_token"].expand(B, -1, -1))
    pos = sd.get("pos_embed")
    if pos is not None and cfg.get("dynamic_img_size", False):                     # :1049-1058
        npt = 0 if cfg.get("no_embed_class", False) else sum(t.shape[1] for t in to_cat)
        old = (cfg["img_size"] // cfg["patch_size"],) * 2
        pos = resample_abs_pos_embed(pos, (img.shape[-2] // cfg["patch_size"], img.shape[-1] // cfg["patch_size"]), old, npt)
    if cfg.get("no_embed_class", False):
        if pos is not None:
            x = x + pos
        x = torch.cat(to_cat + [x], 1)
    else:
        x = torch.cat(to_cat + [x], 1)
        if pos is not None:
            x = x + pos
    if "norm_pre.weight" in sd:
        x = norm(x, sd, "norm_pre.", eps)
    outs = []
    for i in range(cfg["depth"]):
        x = block(x, sd, f"blocks.{i}.", cfg["num_heads"], eps, kv_len, glu=cfg.get("glu"),
                  act="gelu" if cfg.get("mlp_layer") else (cfg.get("act_layer") or "gelu"))
        outs.append(x)
    if "norm.weight" in sd:
        x = norm(x, sd, "norm.", eps)
    return (x, outs) if return_blocks else x


def vit_forward_head(x, sd, cfg):
    """timm/models/vision_transformer.py:1254-1292 (pool :686-708, fc_norm, head)."""
    pool = cfg.get("global_pool", "token")
    npre = cfg.get("num_prefix_tokens", 1)
    if pool == "token":
        x = x[:, 0]
    elif pool == "avg":
        x = x[:, npre:].mean(1)
    elif pool == "map":                                              # :1265-1269 (prefix tokens excluded)
        x = attention_pool_latent(x[:, npre:], sd, "attn_pool.", cfg["num_heads"], cfg.get("eps", 1e-6),
                                  act=cfg.get("act_layer") or "gelu")
    elif pool:
        raise ValueError(f"oracle: pool {pool!r} not restated")
    if "fc_norm.weight" in sd:
        x = norm(x, sd, "fc_norm.", cfg.get("eps", 1e-6))
    if "head.weight" in sd:
        x = linear(x, sd["head.weight"], sd.get("head.bias"))
    return x


def vit_forward(img, sd, cfg, kv_len=None):
    """timm/models/vision_transformer.py:1294-1302."""
    return vit_forward_head(vit_forward_features(img, sd, cfg, kv_len), sd, cfg)


# ----------------------------------------------------------------------------------------------- helpers
VIT_CONFIGS = {
    # name: timm kwargs (the same dict is passed to timm.models.VisionTransformer when making goldens)
    "vit_tiny_patch16_224": dict(img_size=224, patch_size=16, embed_dim=192, depth=12, num_heads=3),
    "vit_base_patch16_224": dict(img_size=224, patch_size=16, embed_dim=768, depth=12, num_heads=12),
    "vit_large_patch16_224": dict(img_size=224, patch_size=16, embed_dim=1024, depth=24, num_heads=16),
}


def random_state_dict(cfg, seed=0, num_classes=1000, dtype=torch.float32):
    """Random weights with timm's ViT key names / shapes (trunc-normal-like std .02 linears, small random
    biases so that bias paths are exercised).  For full-size parity runs where no golden file exists."""
    g = torch.Generator().manual_seed(seed)
    D, p, C = cfg["embed_dim"], cfg["patch_size"], cfg.get("in_chans", 3)
    hid = int(D * cfg.get("mlp_ratio", 4.0))
    G = (cfg["img_size"] // p) ** 2
    npre = (1 if cfg.get("class_token", True) else 0) + cfg.get("reg_tokens", 0)
    rn = lambda *s, std=0.02: (torch.randn(*s, generator=g) * std).to(dtype)  # noqa: E731
    sd = {"patch_embed.proj.weight": rn(D, C, p, p, std=0.05), "patch_embed.proj.bias": rn(D),
          "pos_embed": rn(1, G + npre, D)}
    if cfg.get("class_token", True):
        sd["cls_token"] = rn(1, 1, D)
    if cfg.get("reg_tokens", 0):
        sd["reg_token"] = rn(1, cfg["reg_tokens"], D)
    for i in range(cfg["depth"]):
        pre = f"blocks.{i}."
        sd.update({
            pre + "norm1.weight": 1 + rn(D, std=0.1), pre + "norm1.bias": rn(D, std=0.05),
            pre + "attn.qkv.weight": rn(3 * D, D), pre + "attn.qkv.bias": rn(3 * D),
            pre + "attn.proj.weight": rn(D, D), pre + "attn.proj.bias": rn(D),
            pre + "norm2.weight": 1 + rn(D, std=0.1), pre + "norm2.bias": rn(D, std=0.05),
            pre + "mlp.fc1.weight": rn(hid, D), pre + "mlp.fc1.bias": rn(hid),
            pre + "mlp.fc2.weight": rn(D, hid), pre + "mlp.fc2.bias": rn(D),
        })
        if cfg.get("init_values"):
            sd[pre + "ls1.gamma"] = 0.5 + rn(D, std=0.1)
            sd[pre + "ls2.gamma"] = 0.5 + rn(D, std=0.1)
    sd["norm.weight"] = 1 + rn(D, std=0.1)
    sd["norm.bias"] = rn(D, std=0.05)
    sd["head.weight"] = rn(num_classes, D)
    sd["head.bias"] = rn(num_classes)
    return sd


# ----------------------------------------------------------------------------------------------- timing port
def _eager_norm(x, sd, prefix, eps):
    """The torch library call timm's norm layer makes: F.layer_norm (norm.py:85-90) or, for RmsNorm (no bias in the state
    dict), the pow / mean / rsqrt / mul sequence of fast_norm.py:126-139."""
    import torch.nn.functional as F
    if prefix + "bias" in sd:
        return F.layer_norm(x, (x.shape[-1],), sd[prefix + "weight"], sd[prefix + "bias"], eps)
    v = torch.mean(x.pow(2), dim=(-1,), keepdim=True)
    return x * torch.rsqrt(v + eps) * sd[prefix + "weight"]


def vit_forward_eager(img, sd, cfg, kv_len=None):
    """Same computation as vit_forward, issued as the reference issues it on the CPU -- F.conv2d,
    F.layer_norm, F.linear, F.scaled_dot_product_attention, F.gelu (the call sites listed in SURVEY.md 8c) --
    so that bench.py's cpu_baseline / --impl reference time the library kernels timm's eager path really
    runs on the host, not the slower explicit restatement above.  Pinned to the explicit oracle and to
    the golden vectors by tests/test_oracle_golden.py.  Works in fp32 or bf16 (dtype of `img` / `sd`)."""
    import torch.nn.functional as F
    eps = cfg.get("eps", 1e-6)
    H = cfg["num_heads"]
    act_fn = {"gelu": F.gelu, "gelu_tanh": lambda t: F.gelu(t, approximate="tanh"), "silu": F.silu}[cfg.get("act_layer") or "gelu"]
    x = F.conv2d(img, sd["patch_embed.proj.weight"], sd.get("patch_embed.proj.bias"), stride=cfg["patch_size"])
    x = x.flatten(2).transpose(1, 2)
    B = x.shape[0]
    to_cat = [sd[k].expand(B, -1, -1) for k in ("cls_token", "reg_token") if k in sd]
    pos = sd.get("pos_embed")
    if pos is not None and cfg.get("dynamic_img_size", False):
        npt = 0 if cfg.get("no_embed_class", False) else sum(t.shape[1] for t in to_cat)
        pos = resample_abs_pos_embed(pos, (img.shape[-2] // cfg["patch_size"], img.shape[-1] // cfg["patch_size"]),
                                     (cfg["img_size"] // cfg["patch_size"],) * 2, npt)
    if cfg.get("no_embed_class", False):
        x = x + pos if pos is not None else x
        x = torch.cat(to_cat + [x], 1)
    else:
        x = torch.cat(to_cat + [x], 1)
        x = x + pos if pos is not None else x
    D = x.shape[-1]
    if "norm_pre.weight" in sd:
        x = _eager_norm(x, sd, "norm_pre.", eps)
    mask = None
    if kv_len is not None:
        N = x.shape[1]
        mask = (torch.arange(N)[None, :] < kv_len[:, None])[:, None, None, :]
    for i in range(cfg["depth"]):
        p = f"blocks.{i}."
        h = _eager_norm(x, sd, p + "norm1.", eps)
        N = h.shape[1]
        qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd.get(p + "attn.qkv.bias")).reshape(B, N, 3, H, D // H)
        q, k, v = qkv.permute(2, 0, 3, 1, 4).unbind(0)
        if p + "attn.q_norm.weight" in sd:       # qk_norm, attention.py:117
            q, k = _eager_norm(q, sd, p + "attn.q_norm.", eps), _eager_norm(k, sd, p + "attn.k_norm.", eps)
        h = F.scaled_dot_product_attention(q, k, v, attn_mask=mask).transpose(1, 2).reshape(B, N, D)
        if p + "attn.norm.weight" in sd:         # scale_norm, attention.py:136
            h = _eager_norm(h, sd, p + "attn.norm.", eps)
        h = F.linear(h, sd[p + "attn.proj.weight"], sd.get(p + "attn.proj.bias"))
        if p + "ls1.gamma" in sd:
            h = h * sd[p + "ls1.gamma"]
        x = x + h
        h = _eager_norm(x, sd, p + "norm2.", eps)
        if p + "mlp.fc1_g.weight" in sd:        # SwiGLU.forward, timm/layers/mlp.py:153-162
            h = F.silu(F.linear(h, sd[p + "mlp.fc1_g.weight"], sd.get(p + "mlp.fc1_g.bias"))) * \
                F.linear(h, sd[p + "mlp.fc1_x.weight"], sd.get(p + "mlp.fc1_x.bias"))
        else:
            h = act_fn(F.linear(h, sd[p + "mlp.fc1.weight"], sd.get(p + "mlp.fc1.bias")))
        h = F.linear(h, sd[p + "mlp.fc2.weight"], sd.get(p + "mlp.fc2.bias"))
        if p + "ls2.gamma" in sd:
            h = h * sd[p + "ls2.gamma"]
        x = x + h
    if "norm.weight" in sd:
        x = _eager_norm(x, sd, "norm.", eps)
    pool = cfg.get("global_pool", "token")
    if pool == "token":
        x = x[:, 0]
    elif pool == "avg":
        x = x[:, cfg.get("num_prefix_tokens", 1):].mean(1)
    elif pool == "map":      # AttentionPoolLatent.forward, attention_pool.py:95-128, as its eager call sequence
        x = x[:, cfg.get("num_prefix_tokens", 1):]
        Bp, Np, C = x.shape
        q = F.linear(sd["attn_pool.latent"].expand(Bp, -1, -1), sd["attn_pool.q.weight"], sd.get("attn_pool.q.bias"))
        q = q.reshape(Bp, 1, H, C // H).transpose(1, 2)
        kv = F.linear(x, sd["attn_pool.kv.weight"], sd.get("attn_pool.kv.bias")).reshape(Bp, Np, 2, H, C // H)
        k, v = kv.permute(2, 0, 3, 1, 4).unbind(0)
        x = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(Bp, 1, C)
        x = F.linear(x, sd["attn_pool.proj.weight"], sd.get("attn_pool.proj.bias"))
        h = _eager_norm(x, sd, "attn_pool.norm.", eps) if "attn_pool.norm.weight" in sd else x
        h = F.linear(act_fn(F.linear(h, sd["attn_pool.mlp.fc1.weight"], sd.get("attn_pool.mlp.fc1.bias"))),
                     sd["attn_pool.mlp.fc2.weight"], sd.get("attn_pool.mlp.fc2.bias"))
        x = (x + h)[:, 0]
    if "fc_norm.weight" in sd:
        x = _eager_norm(x, sd, "fc_norm.", eps)
    if "head.weight" in sd:
        x = F.linear(x, sd["head.weight"], sd.get("head.bias"))
    return x


def normalize_u8(img_u8, mean, std):
    """PrefetchLoader's device-side normalisation, timm/data/loader.py:106-109,139: mean / std (per channel, [0, 1] units)
    become fp32 tensors of `x * 255`; the uint8 batch is widened to fp32, the mean subtracted and the result divided by std.
    numpy float32 arithmetic (IEEE subtract, then divide) -- what the CUDA path's per-channel lookup table reproduces."""
    import numpy as np
    x = img_u8.numpy().astype(np.float32)
    m = np.asarray([v * 255 for v in mean], dtype=np.float32).reshape(1, -1, 1, 1)
    s = np.asarray([v * 255 for v in std], dtype=np.float32).reshape(1, -1, 1, 1)
    return torch.from_numpy((x - m) / s)
