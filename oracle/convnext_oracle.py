"""ORACLE (test infrastructure, NOT product code): CPU fp32 restatement of timm's ConvNeXt forward.

Same rules as oracle/vit_oracle.py: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU baseline.
Pinned against golden vectors made from the unmodified reference (oracle/make_golden.py ->
tests/golden/convnext_g1.pt) by tests/test_oracle_golden.py.  Each function cites the reference lines it restates
(timm 1.0.28.dev0)."""
import torch

from .vit_oracle import ACTS, layer_norm, linear


def dwconv(x, weight, bias):
    """nn.Conv2d(C, C, k, padding=k//2, groups=C) on NCHW (timm/models/convnext.py:173-182): explicit shifted sums."""
    k = weight.shape[-1]
    p = k // 2
    xp = torch.nn.functional.pad(x, (p, p, p, p))
    H, W = x.shape[-2:]
    y = torch.zeros_like(x)
    for ky in range(k):
        for kx in range(k):
            y = y + xp[:, :, ky:ky + H, kx:kx + W] * weight[:, 0, ky, kx].view(1, -1, 1, 1)
    return y if bias is None else y + bias.view(1, -1, 1, 1)


def layer_norm2d(x, weight, bias, eps):
    """timm/layers/norm.py:127-134 (LayerNorm2d): LayerNorm over C of NCHW."""
    return layer_norm(x.permute(0, 2, 3, 1), weight, bias, eps).permute(0, 3, 1, 2)


def patch_conv(x, weight, bias):
    """Conv2d(kernel = stride = p): stem (convnext.py:408-411) and downsample (:263-275) convolutions."""
    p = weight.shape[-1]
    B, C, H, W = x.shape
    g = x.reshape(B, C, H // p, p, W // p, p).permute(0, 2, 4, 1, 3, 5).reshape(B, H // p, W // p, C * p * p)
    return linear(g, weight.reshape(weight.shape[0], -1), bias).permute(0, 3, 1, 2)


def block(x, sd, pre, eps, keep=None, keep_prob=1.0, act="gelu"):
    """ConvNeXtBlock.forward, timm/models/convnext.py:197-213 (conv_mlp=False).  keep: optional per-sample DropPath
    keep mask [B] (1 = kept), applied as timm/layers/drop.py:158-175 does: branch * mask / keep_prob."""
    h = dwconv(x, sd[pre + "conv_dw.weight"], sd.get(pre + "conv_dw.bias")).permute(0, 2, 3, 1)      # :200,205
    h = layer_norm(h, sd[pre + "norm.weight"], sd[pre + "norm.bias"], eps)                             # :206
    h = ACTS[act](linear(h, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"]))        # act_layer: create_act.py key
    if pre + "mlp.grn.weight" in sd:     # ConvNeXt-V2: GlobalResponseNorm, timm/layers/grn.py:46-49 (eps 1e-6), mlp.py:289-292
        gx = torch.sqrt((h * h).sum(dim=(1, 2), keepdim=True))
        nx = gx / (gx.mean(dim=-1, keepdim=True) + 1e-6)
        h = h + sd[pre + "mlp.grn.bias"] + sd[pre + "mlp.grn.weight"] * (h * nx)
    h = linear(h, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"]).permute(0, 3, 1, 2)           # :207-208
    if pre + "gamma" in sd:
        h = h * sd[pre + "gamma"].view(1, -1, 1, 1)                                                    # :209-210
    if keep is not None:
        h = h * (keep.view(-1, 1, 1, 1) / keep_prob)                                                   # :213 drop_path
    return h + x                                                                                       # :212


def convnext_forward_features(img, sd, cfg, return_stages=False, keeps=None):
    """ConvNeXt.forward_features, timm/models/convnext.py:606-611.  keeps: optional {block prefix: (mask [B], keep_prob)}
    of DropPath draws (training mode)."""
    eps = cfg.get("eps", 1e-6)
    x = layer_norm2d(patch_conv(img, sd["stem.0.weight"], sd.get("stem.0.bias")), sd["stem.1.weight"], sd["stem.1.bias"], eps)
    outs = []
    for i, depth in enumerate(cfg["depths"]):
        pre = f"stages.{i}."
        if pre + "downsample.1.weight" in sd:                                                          # :300-302
            x = layer_norm2d(x, sd[pre + "downsample.0.weight"], sd[pre + "downsample.0.bias"], eps)
            x = patch_conv(x, sd[pre + "downsample.1.weight"], sd.get(pre + "downsample.1.bias"))
        for j in range(depth):
            kp = (keeps or {}).get(f"{pre}blocks.{j}.")
            x = block(x, sd, f"{pre}blocks.{j}.", eps, *(kp if kp is not None else (None, 1.0)), act=cfg.get("act_layer") or "gelu")
        outs.append(x)
    return (x, outs) if return_stages else x


def convnext_forward_head(x, sd, cfg):
    """NormMlpClassifierHead.forward, timm/layers/classifier.py:211-220: avg pool -> LayerNorm2d -> flatten -> fc."""
    x = x.mean((2, 3))
    x = layer_norm(x, sd["head.norm.weight"], sd["head.norm.bias"], cfg.get("eps", 1e-6))
    return linear(x, sd["head.fc.weight"], sd.get("head.fc.bias"))


def convnext_forward(img, sd, cfg):
    return convnext_forward_head(convnext_forward_features(img, sd, cfg), sd, cfg)


def convnext_forward_eager(img, sd, cfg):
    """The same computation issued as timm's eager path issues it (F.conv2d / F.layer_norm / F.linear / F.gelu):
    the CPU baseline bench.py times for the ConvNeXt workload."""
    import torch.nn.functional as F
    eps = cfg.get("eps", 1e-6)
    act_fn = {"gelu": F.gelu, "gelu_tanh": lambda t: F.gelu(t, approximate="tanh"), "silu": F.silu}[cfg.get("act_layer") or "gelu"]

    def ln2d(t, w, b):
        return F.layer_norm(t.permute(0, 2, 3, 1), (t.shape[1],), w, b, eps).permute(0, 3, 1, 2)

    p = sd["stem.0.weight"].shape[-1]
    x = ln2d(F.conv2d(img, sd["stem.0.weight"], sd.get("stem.0.bias"), stride=p), sd["stem.1.weight"], sd["stem.1.bias"])
    for i, depth in enumerate(cfg["depths"]):
        pre = f"stages.{i}."
        if pre + "downsample.1.weight" in sd:
            x = ln2d(x, sd[pre + "downsample.0.weight"], sd[pre + "downsample.0.bias"])
            x = F.conv2d(x, sd[pre + "downsample.1.weight"], sd.get(pre + "downsample.1.bias"), stride=2)
        for j in range(depth):
            b = f"{pre}blocks.{j}."
            k = sd[b + "conv_dw.weight"].shape[-1]
            h = F.conv2d(x, sd[b + "conv_dw.weight"], sd.get(b + "conv_dw.bias"), padding=k // 2, groups=x.shape[1])
            h = F.layer_norm(h.permute(0, 2, 3, 1), (x.shape[1],), sd[b + "norm.weight"], sd[b + "norm.bias"], eps)
            h = act_fn(F.linear(h, sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"]))
            if b + "mlp.grn.weight" in sd:      # GlobalResponseNorm.forward, timm/layers/grn.py:46-49
                x_g = h.norm(p=2, dim=(1, 2), keepdim=True)
                x_n = x_g / (x_g.mean(dim=-1, keepdim=True) + 1e-6)
                h = h + torch.addcmul(sd[b + "mlp.grn.bias"].view(1, 1, 1, -1), sd[b + "mlp.grn.weight"].view(1, 1, 1, -1), h * x_n)
            h = F.linear(h, sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"]).permute(0, 3, 1, 2)
            if b + "gamma" in sd:
                h = h * sd[b + "gamma"].view(1, -1, 1, 1)
            x = h + x
    x = F.layer_norm(x.mean((2, 3)), (x.shape[1],), sd["head.norm.weight"], sd["head.norm.bias"], eps)
    return F.linear(x, sd["head.fc.weight"], sd.get("head.fc.bias"))


CONVNEXT_CONFIGS = {
    "convnext_base": dict(depths=(3, 3, 27, 3), dims=(128, 256, 512, 1024)),
}


def random_state_dict(cfg, seed=0, num_classes=1000, in_chans=3, patch=4, ksize=7, ls=0.5):
    """Random weights with timm's ConvNeXt key names / shapes (LayerScale gamma ~0.5 so the branch matters)."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=0.02: torch.randn(*s, generator=g) * std  # noqa: E731
    dims, depths = cfg["dims"], cfg["depths"]
    sd = {"stem.0.weight": rn(dims[0], in_chans, patch, patch, std=0.1), "stem.0.bias": rn(dims[0]),
          "stem.1.weight": 1 + rn(dims[0], std=0.1), "stem.1.bias": rn(dims[0], std=0.05)}
    prev = dims[0]
    for i, (d, depth) in enumerate(zip(dims, depths)):
        pre = f"stages.{i}."
        if i > 0:
            sd.update({pre + "downsample.0.weight": 1 + rn(prev, std=0.1), pre + "downsample.0.bias": rn(prev, std=0.05),
                       pre + "downsample.1.weight": rn(d, prev, 2, 2, std=0.05), pre + "downsample.1.bias": rn(d)})
        for j in range(depth):
            b = f"{pre}blocks.{j}."
            sd.update({b + "gamma": ls + rn(d, std=0.1), b + "conv_dw.weight": rn(d, 1, ksize, ksize, std=0.1),
                       b + "conv_dw.bias": rn(d), b + "norm.weight": 1 + rn(d, std=0.1), b + "norm.bias": rn(d, std=0.05),
                       b + "mlp.fc1.weight": rn(4 * d, d, std=0.04), b + "mlp.fc1.bias": rn(4 * d),
                       b + "mlp.fc2.weight": rn(d, 4 * d, std=0.04), b + "mlp.fc2.bias": rn(d)})
        prev = d
    sd.update({"head.norm.weight": 1 + rn(prev, std=0.1), "head.norm.bias": rn(prev, std=0.05),
               "head.fc.weight": rn(num_classes, prev), "head.fc.bias": rn(num_classes)})
    return sd
