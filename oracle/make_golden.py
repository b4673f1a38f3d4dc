"""Generate tests/golden/*.pt from the UNMODIFIED reference (timm at /root/reference), CPU fp32.

Run in the authoring container only (the GPU box has no /root/reference):
    python oracle/make_golden.py
Each fixture holds the reference module's state_dict, the seeded inputs and the reference outputs /
gradients.  The oracle restatement (oracle/vit_oracle.py) is pinned against them by
tests/test_oracle_golden.py; the CUDA path is checked against them by tests/test_gpu_models.py.
Head dim is 64 in every config (the tcgen05 attention kernel's supported head size).
"""
import os
import sys

import torch

REF = os.environ.get("TIMM_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _half_sd(sd):
    return {k: v.clone() for k, v in sd.items()}


def vit_case(name, seed, batch, kwargs, backward=True, input_size=None):
    import timm  # noqa: F401
    from timm.models.vision_transformer import VisionTransformer
    torch.manual_seed(seed)
    build = dict(kwargs)
    if isinstance(build.get("mlp_layer"), str):     # fixtures keep names; timm's constructor takes the classes
        import timm.layers
        build["mlp_layer"] = getattr(timm.layers, build["mlp_layer"])
    if isinstance(build.get("act_layer"), str):
        from timm.layers import get_act_layer
        build["act_layer"] = get_act_layer(build["act_layer"])
    model = VisionTransformer(**build)
    # default init zeroes biases / sets LN to (1, 0): perturb so every parameter path is exercised
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
            if n in ("cls_token", "reg_token", "pos_embed"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    model.eval()
    ih, iw = input_size or (kwargs["img_size"], kwargs["img_size"])      # dynamic_img_size models: not the nominal size
    img = torch.randn(batch, kwargs.get("in_chans", 3), ih, iw, generator=g)
    feats_per_block = []
    hooks = [blk.register_forward_hook(lambda m, i, o: feats_per_block.append(o.detach().clone()))
             for blk in model.blocks]
    feats = model.forward_features(img)
    logits = model.forward_head(feats)
    for h in hooks:
        h.remove()
    fx = {"kind": "vit", "kwargs": kwargs, "state_dict": _half_sd(model.state_dict()), "img": img,
          "features": feats.detach(), "logits": logits.detach(), "block_outputs": feats_per_block,
          "timm_version": timm.__version__, "torch_version": str(torch.__version__)}
    if backward:
        # the reference's own backward test pattern: outputs.mean().backward()  (tests/test_models.py:209-242)
        model.zero_grad()
        model(img).mean().backward()
        fx["grads"] = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    torch.save(fx, os.path.join(OUT, name + ".pt"))
    print(name, "logits", tuple(logits.shape), "absmax", logits.abs().max().item(),
          "params", sum(p.numel() for p in model.parameters()))


def convnext_case(name, seed, batch, kwargs, img_size, backward=False):
    import timm
    from timm.models.convnext import ConvNeXt
    torch.manual_seed(seed)
    model = ConvNeXt(**kwargs)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:   # biases, norm weights, LayerScale gamma (1e-6 by default would hide the branch)
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
            if n.endswith("gamma"):
                p.add_(0.5)
    model.eval()
    img = torch.randn(batch, 3, img_size, img_size, generator=g)
    stage_out = []
    hooks = [s.register_forward_hook(lambda m, i, o: stage_out.append(o.detach().clone())) for s in model.stages]
    feats = model.forward_features(img)
    logits = model.forward_head(feats)
    for h in hooks:
        h.remove()
    fx = {"kind": "convnext", "kwargs": kwargs, "state_dict": _half_sd(model.state_dict()), "img": img,
          "features": feats.detach(), "logits": logits.detach(), "stage_outputs": stage_out,
          "timm_version": timm.__version__, "torch_version": str(torch.__version__)}
    if backward:        # the reference's backward test pattern (tests/test_models.py:209-242)
        model.zero_grad()
        model(img).mean().backward()
        fx["grads"] = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    torch.save(fx, os.path.join(OUT, name + ".pt"))
    print(name, "logits", tuple(logits.shape), "absmax", logits.abs().max().item(),
          "params", sum(p.numel() for p in model.parameters()))


def naflex_batch(g, batch, max_grid, patch, chans=3):
    """A NaFlex batch the way NaFlexCollator lays it out (timm/data/naflex_dataset.py:122-150): per-sample (h, w) grid,
    row-major coords, valid prefix, zero padding up to max_grid[0]*max_grid[1] tokens."""
    N = max_grid[0] * max_grid[1]
    grids = [max_grid, (3, 5), (4, 2), (2, 2), (5, 3)][:batch]
    patches = torch.zeros(batch, N, patch * patch * chans)
    coord = torch.zeros(batch, N, 2, dtype=torch.int64)
    valid = torch.zeros(batch, N, dtype=torch.bool)
    for b, (h, w) in enumerate(grids):
        n = h * w
        patches[b, :n] = torch.randn(n, patch * patch * chans, generator=g)
        yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        coord[b, :n] = torch.stack([yy.flatten(), xx.flatten()], -1)
        valid[b, :n] = True
    return patches, coord, valid


def naflex_case(name, seed, batch):
    import timm
    from timm.models.naflexvit import NaFlexVit, NaFlexVitCfg
    torch.manual_seed(seed)
    cfg = dict(patch_size=14, embed_dim=128, depth=2, num_heads=2, class_token=True, pos_embed_grid_size=(6, 6),
               pre_norm=True, global_pool="token")
    model = NaFlexVit(cfg=NaFlexVitCfg(**cfg), in_chans=3, num_classes=9, img_size=84)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
            if n in ("embeds.cls_token", "embeds.pos_embed"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    model.eval()
    patches, coord, valid = naflex_batch(g, batch, (6, 6), 14)
    with torch.no_grad():
        feats = model.forward_features(patches, patch_coord=coord, patch_valid=valid)
        logits = model.forward_head(feats["patches"], patch_valid=valid)
    fx = {"kind": "naflex", "kwargs": dict(cfg, num_classes=9), "state_dict": _half_sd(model.state_dict()),
          "patches": patches, "patch_coord": coord, "patch_valid": valid, "features": feats["patches"].detach(),
          "logits": logits.detach(), "timm_version": timm.__version__, "torch_version": str(torch.__version__)}
    # backward the way the reference's own model tests do it (tests/test_models.py:209-242): outputs.mean().backward()
    model.zero_grad()
    model(patches, patch_coord=coord, patch_valid=valid).mean().backward()
    fx["grads"] = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    torch.save(fx, os.path.join(OUT, name + ".pt"))
    print(name, "logits", tuple(logits.shape), "absmax", logits.abs().max().item())


def layer_cases():
    from timm.layers import Attention, LayerNorm, Mlp, PatchEmbed, LayerScale, DropPath
    from timm.models.naflexvit import create_attention_mask
    g = torch.Generator().manual_seed(7)
    fx = {"kind": "layers"}
    torch.manual_seed(7)
    # Attention with and without a NaFlex-style key-padding mask
    attn = Attention(128, num_heads=2, qkv_bias=True).eval()
    with torch.no_grad():
        attn.qkv.bias.add_(torch.randn(attn.qkv.bias.shape, generator=g) * 0.1)
        attn.proj.bias.add_(torch.randn(attn.proj.bias.shape, generator=g) * 0.1)
    x = torch.randn(3, 37, 128, generator=g)
    valid = torch.zeros(3, 37, dtype=torch.bool)
    lens = [37, 20, 5]
    for i, n in enumerate(lens):
        valid[i, :n] = True
    mask = create_attention_mask(valid, num_prefix_tokens=0, dtype=torch.float32)
    fx["attention"] = {"state_dict": _half_sd(attn.state_dict()), "x": x, "num_heads": 2, "y": attn(x).detach(),
                       "kv_len": torch.tensor(lens, dtype=torch.int32), "y_masked": attn(x, attn_mask=mask).detach()}
    mlp = Mlp(128, 512).eval()
    fx["mlp"] = {"state_dict": _half_sd(mlp.state_dict()), "x": x, "y": mlp(x).detach()}
    ln = LayerNorm(128, eps=1e-6)
    with torch.no_grad():
        ln.weight.add_(torch.randn(128, generator=g) * 0.2)
        ln.bias.add_(torch.randn(128, generator=g) * 0.2)
    fx["layernorm"] = {"state_dict": _half_sd(ln.state_dict()), "x": x * 3 + 1, "eps": 1e-6, "y": ln(x * 3 + 1).detach()}
    for p, size in ((16, 64), (14, 56), (4, 32)):
        pe = PatchEmbed(img_size=size, patch_size=p, in_chans=3, embed_dim=64).eval()
        img = torch.randn(2, 3, size, size, generator=g)
        fx[f"patch_embed_p{p}"] = {"state_dict": _half_sd(pe.state_dict()), "img": img, "patch": p, "y": pe(img).detach()}
    ls = LayerScale(128, init_values=0.3)
    fx["layerscale"] = {"state_dict": _half_sd(ls.state_dict()), "x": x, "y": ls(x).detach()}
    # DropPath semantics (timm/layers/drop.py:158-175): kept samples scaled by exactly 1/keep_prob
    torch.manual_seed(3)
    dp = DropPath(0.25).train()
    xin = torch.ones(64, 4, 8)
    y = dp(xin)
    fx["droppath"] = {"drop_prob": 0.25, "x": xin, "y": y, "keep_mask": (y[:, 0, 0] != 0).to(torch.int32)}
    torch.save(fx, os.path.join(OUT, "layers.pt"))
    print("layers fixture written")


def rope_cases():
    """timm.layers.AttentionRope with the [sin | cos] table of RotaryEmbeddingCat (pos_embed_sincos.py:534-715), both
    rotation layouts, one cls + one register token left un-rotated; outputs and all gradients after out.sum()-style
    weighting.  Stored separately (tests/golden/rope.pt) so the other fixtures stay byte-identical."""
    from timm.layers import AttentionRope, RotaryEmbeddingCat
    g = torch.Generator().manual_seed(11)
    fx = {"timm_version": __import__("timm").__version__, "torch_version": str(torch.__version__)}
    dim, heads, grid, npt, B = 128, 2, (6, 5), 2, 3
    N = grid[0] * grid[1] + npt
    rope = RotaryEmbeddingCat(dim // heads, in_pixels=False, feat_shape=list(grid), ref_feat_shape=list(grid))
    emb = rope.get_embed()
    for half in (False, True):
        torch.manual_seed(20 + int(half))
        m = AttentionRope(dim, num_heads=heads, qkv_bias=True, num_prefix_tokens=npt, rotate_half=half).eval()
        x = torch.randn(B, N, dim, generator=g).requires_grad_(True)
        wgt = torch.randn(B, N, dim, generator=g)
        y = m(x, rope=emb)
        (y * wgt).sum().backward()
        fx["half" if half else "interleaved"] = {
            "state_dict": _half_sd(m.state_dict()), "x": x.detach().clone(), "emb": emb.detach().clone(), "y": y.detach().clone(),
            "wgt": wgt, "dx": x.grad.clone(), "grads": {k: p.grad.clone() for k, p in m.named_parameters()},
            "num_heads": heads, "num_prefix": npt, "grid": grid}
    # EvaAttention (timm/models/eva.py:105-262): q_bias / v_bias parameters + zero k_bias buffer, fused qkv, no norms
    from timm.models.eva import EvaAttention
    torch.manual_seed(30)
    m = EvaAttention(dim, num_heads=heads, qkv_bias=True, num_prefix_tokens=npt, scale_norm=False).eval()
    with torch.no_grad():
        m.q_bias.normal_(std=0.5)
        m.v_bias.normal_(std=0.5)
    x = torch.randn(B, N, dim, generator=g).requires_grad_(True)
    wgt = torch.randn(B, N, dim, generator=g)
    y = m(x, rope=emb)
    (y * wgt).sum().backward()
    fx["eva"] = {"state_dict": _half_sd(m.state_dict()), "x": x.detach().clone(), "emb": emb.detach().clone(),
                 "y": y.detach().clone(), "wgt": wgt, "dx": x.grad.clone(),
                 "grads": {k: p.grad.clone() for k, p in m.named_parameters()}, "num_heads": heads, "num_prefix": npt}
    torch.save(fx, os.path.join(OUT, "rope.pt"))
    print("rope fixture written", tuple(emb.shape))


# RmsNorm blocks / final norm (timm/layers/norm.py:202-248), the norm_layer string create_norm._NORM_MAP resolves
RMS_KWARGS = dict(img_size=48, patch_size=16, in_chans=3, num_classes=6, embed_dim=128, depth=2, num_heads=2, qkv_bias=True,
                  norm_layer="rmsnorm")


OPTIM_SHAPES = [(37,), (64, 33), (1000,), (3, 7, 11), (4096 + 5,), (1,), (32, 16)]
OPTIM_CASES = {     # name -> (timm optimizer name, create_optimizer_v2 kwargs, clip_grad max-norm or None)
    "sgd_nesterov_wd": ("sgd", dict(lr=0.1, momentum=0.9, weight_decay=1e-4), None),
    "momentum_dampening": ("momentum", dict(lr=0.05, momentum=0.8, dampening=0.1, weight_decay=0.0), None),
    "sgd_plain": ("momentum", dict(lr=0.2, momentum=0.0, weight_decay=1e-3), None),
    "adamw": ("adamw", dict(lr=1e-3, weight_decay=0.05), None),
    "adam_l2": ("adam", dict(lr=2e-3, weight_decay=1e-4, betas=(0.8, 0.99), eps=1e-6), None),
    "adamw_clip": ("adamw", dict(lr=1e-3, weight_decay=0.05), 1.0),
    "sgd_clip": ("sgd", dict(lr=0.1, momentum=0.9, weight_decay=1e-4), 0.5),
}


def optim_cases(steps=3):
    """Parameter trajectories of the optimizers timm's factory builds for this path (create_optimizer_v2 on a bare
    parameter list: one group, weight decay on every tensor) and of timm.utils.dispatch_clip_grad(mode='norm')."""
    import timm
    from timm.optim import create_optimizer_v2
    from timm.utils import dispatch_clip_grad
    g = torch.Generator().manual_seed(11)
    init = [torch.randn(s, generator=g) for s in OPTIM_SHAPES]
    grads = [[torch.randn(s, generator=g) * (0.5 + k) for s in OPTIM_SHAPES] for k in range(steps)]
    fx = {"kind": "optim", "init": init, "grads": grads, "cases": {}, "timm_version": timm.__version__,
          "torch_version": str(torch.__version__)}
    for name, (opt, kw, clip) in OPTIM_CASES.items():
        params = [torch.nn.Parameter(t.clone()) for t in init]
        o = create_optimizer_v2(params, opt, **kw)
        traj, norms = [], []
        for k in range(steps):
            for p, gk in zip(params, grads[k]):
                p.grad = gk.clone()
            if clip is not None:
                norms.append(float(torch.nn.utils.get_total_norm([p.grad for p in params])))
                dispatch_clip_grad(params, clip, mode="norm")
            o.step()
            traj.append([p.detach().clone() for p in params])
        fx["cases"][name] = {"opt": opt, "kwargs": kw, "clip": clip, "params": traj, "norms": norms,
                             "opt_class": type(o).__name__, "defaults": {k: v for k, v in o.defaults.items()
                                                                         if isinstance(v, (int, float, bool, tuple))}}
        print(name, type(o).__name__, "final |p|", float(sum(p.detach().norm() for p in params)))
    torch.save(fx, os.path.join(OUT, "optim.pt"))


def ema_case(steps=5):
    """timm.utils.ModelEmaV3 (train.py --model-ema --model-ema-warmup) over a drifting model with a BatchNorm buffer
    set: the averaged state_dict after every update, and the decay schedule."""
    import timm
    from timm.utils import ModelEmaV3
    torch.manual_seed(21)
    model = torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.BatchNorm1d(64), torch.nn.Linear(64, 5))
    kw = dict(decay=0.998, use_warmup=True, update_after_step=1, warmup_gamma=1.0, warmup_power=2 / 3)
    ema = ModelEmaV3(model, **kw)
    g = torch.Generator().manual_seed(22)
    fx = {"kind": "ema", "kwargs": kw, "init": {k: v.clone() for k, v in model.state_dict().items()}, "model_states": [],
          "ema_states": [], "decays": [], "timm_version": timm.__version__}
    for step in range(1, steps + 1):
        with torch.no_grad():
            for p in model.parameters():
                p.add_(torch.randn(p.shape, generator=g) * 0.1)
            model[1].running_mean.add_(torch.randn(64, generator=g) * 0.1)
            model[1].num_batches_tracked.add_(1)
        ema.update(model, step=step)
        fx["model_states"].append({k: v.clone() for k, v in model.state_dict().items()})
        fx["ema_states"].append({k: v.clone() for k, v in ema.module.state_dict().items()})
        fx["decays"].append(ema.get_decay(step))
    torch.save(fx, os.path.join(OUT, "ema.pt"))
    print("ema decays", fx["decays"])


def loss_cases():
    """nn.CrossEntropyLoss (train.py:855), timm.loss.LabelSmoothingCrossEntropy, timm.loss.SoftTargetCrossEntropy on
    seeded logits: loss value and d loss / d logits, appended to tests/golden/ema.pt's sibling loss.pt."""
    import timm
    from timm.loss import LabelSmoothingCrossEntropy, SoftTargetCrossEntropy
    g = torch.Generator().manual_seed(31)
    fx = {"kind": "loss", "cases": {}, "timm_version": timm.__version__}
    for name, B, C in (("small", 7, 10), ("imagenet", 16, 1000), ("wide", 3, 2500)):
        x = (torch.randn(B, C, generator=g) * 3).requires_grad_(True)
        t = torch.randint(0, C, (B,), generator=g)
        soft = torch.softmax(torch.randn(B, C, generator=g) * 2, -1)
        case = {"logits": x.detach().clone(), "target": t, "soft_target": soft}
        for key, fn, tgt in (("ce", torch.nn.CrossEntropyLoss(), t), ("ls", LabelSmoothingCrossEntropy(0.1), t),
                             ("soft", SoftTargetCrossEntropy(), soft)):
            x.grad = None
            loss = fn(x, tgt)
            (loss * 1.7).backward()
            case[key] = {"loss": loss.detach().clone(), "dlogits": x.grad.clone(), "grad_out": 1.7}
        fx["cases"][name] = case
    torch.save(fx, os.path.join(OUT, "loss.pt"))
    print("loss cases", {k: float(v["ls"]["loss"]) for k, v in fx["cases"].items()})


def glu_cases():
    """timm.layers.SwiGLU (with and without the inner LayerNorm, EVA-02's scale_mlp), SwiGLUPacked and the default GluMlp
    (sigmoid gate on the second half): output, input gradient and parameter gradients -> tests/golden/glu.pt; and a ViT
    with mlp_layer=SwiGLU (logits + parameter gradients) -> tests/golden/vit_g5_swiglu.pt."""
    import timm
    from timm.layers import GluMlp, LayerNorm, SwiGLU, SwiGLUPacked
    g = torch.Generator().manual_seed(41)
    fx = {"kind": "glu", "timm_version": timm.__version__}
    torch.manual_seed(41)
    mods = {"swiglu": SwiGLU(64, 128), "swiglu_norm": SwiGLU(64, 128, norm_layer=LayerNorm),
            "swiglu_packed": SwiGLUPacked(64, 256), "glu_sigmoid": GluMlp(64, 256, bias=(True, False))}
    for name, m in mods.items():
        with torch.no_grad():
            for p in m.parameters():
                if p.dim() == 1:
                    p.add_(torch.randn(p.shape, generator=g) * 0.05)
        x = torch.randn(3, 50, 64, generator=g, requires_grad=True)
        wgt = torch.randn(3, 50, 64, generator=g)
        y = m(x)
        (y * wgt).sum().backward()
        fx[name] = {"state_dict": {k: v.clone() for k, v in m.state_dict().items()}, "x": x.detach().clone(), "y": y.detach(),
                    "wgt": wgt, "dx": x.grad.clone(), "grads": {k: p.grad.clone() for k, p in m.named_parameters()},
                    "act": type(m.act).__name__, "gate_last": getattr(m, "gate_last", None)}
    torch.save(fx, os.path.join(OUT, "glu.pt"))
    print("glu cases", {k: float(v["y"].abs().max()) for k, v in fx.items() if isinstance(v, dict)})
    vit_case("vit_g5_swiglu", 6, 2, dict(img_size=64, patch_size=16, in_chans=3, num_classes=10, embed_dim=128, depth=2,
                                         num_heads=2, mlp_ratio=2.0, qkv_bias=True, init_values=0.1,
                                         mlp_layer="SwiGLU", act_layer="silu"))


def qknorm_cases():
    """timm.layers.Attention with qk_norm (LayerNorm and RmsNorm over head_dim) and scale_norm: output, input gradient and
    parameter gradients, with and without key padding -> tests/golden/qknorm.pt; and a ViT with qk_norm=True,
    scale_attn_norm=True -> tests/golden/vit_g6_qknorm.pt."""
    import timm
    from timm.layers import Attention, LayerNorm, RmsNorm
    g = torch.Generator().manual_seed(51)
    torch.manual_seed(51)
    fx = {"kind": "qknorm", "timm_version": timm.__version__}
    mods = {"ln_qk": Attention(128, num_heads=2, qkv_bias=True, qk_norm=True, norm_layer=LayerNorm),
            "rms_qk_scale": Attention(128, num_heads=2, qkv_bias=False, qk_norm=True, scale_norm=True, norm_layer=RmsNorm),
            "ln_scale_only": Attention(128, num_heads=2, qkv_bias=True, scale_norm=True, norm_layer=LayerNorm)}
    for name, m in mods.items():
        with torch.no_grad():
            for p in m.parameters():
                if p.dim() == 1:
                    p.add_(torch.randn(p.shape, generator=g) * 0.1)
        x = torch.randn(3, 37, 128, generator=g, requires_grad=True)
        wgt = torch.randn(3, 37, 128, generator=g)
        y = m(x)
        (y * wgt).sum().backward()
        fx[name] = {"state_dict": {k: v.clone() for k, v in m.state_dict().items()}, "x": x.detach().clone(), "y": y.detach(),
                    "wgt": wgt, "dx": x.grad.clone(), "grads": {k: p.grad.clone() for k, p in m.named_parameters()},
                    "num_heads": 2, "norm": type(m.q_norm if name != "ln_scale_only" else m.norm).__name__}
    torch.save(fx, os.path.join(OUT, "qknorm.pt"))
    print("qknorm cases", {k: float(v["y"].abs().max()) for k, v in fx.items() if isinstance(v, dict)})
    vit_case("vit_g6_qknorm", 7, 2, dict(img_size=64, patch_size=16, in_chans=3, num_classes=10, embed_dim=128, depth=2,
                                         num_heads=2, mlp_ratio=2.0, qkv_bias=False, qk_norm=True, scale_attn_norm=True))


DYN_KWARGS = dict(img_size=64, patch_size=16, in_chans=3, num_classes=10, embed_dim=128, depth=1, num_heads=2, mlp_ratio=2.0,
                  qkv_bias=True, dynamic_img_size=True, reg_tokens=1)


SIGLIP_KWARGS = dict(img_size=64, patch_size=16, in_chans=3, num_classes=10, embed_dim=128, depth=1, num_heads=2, mlp_ratio=2.0,
                     qkv_bias=True, class_token=False, global_pool="map", act_layer="gelu_tanh")   # SigLIP: tanh GELU + map head
CNX_TANH_KWARGS = dict(depths=(1, 1, 1, 1), dims=(16, 32, 64, 128), num_classes=10, kernel_sizes=(7, 5, 5, 3),
                       act_layer="gelu_tanh")                                                      # timm's test_convnext family


def mlp_act_cases():
    """timm.layers.Mlp with act_layer 'gelu_tanh' / 'silu' (create_act.py): output, input gradient, parameter gradients."""
    import timm
    from timm.layers import Mlp, get_act_layer
    g = torch.Generator().manual_seed(61)
    torch.manual_seed(61)
    fx = {"kind": "mlp_acts", "timm_version": timm.__version__}
    for name in ("gelu_tanh", "silu"):
        m = Mlp(64, 256, act_layer=get_act_layer(name))
        with torch.no_grad():
            for p in m.parameters():
                if p.dim() == 1:
                    p.add_(torch.randn(p.shape, generator=g) * 0.1)
        x = (torch.randn(3, 50, 64, generator=g) * 2).requires_grad_(True)
        wgt = torch.randn(3, 50, 64, generator=g)
        y = m(x)
        (y * wgt).sum().backward()
        fx[name] = {"state_dict": {k: v.clone() for k, v in m.state_dict().items()}, "x": x.detach().clone(), "y": y.detach(),
                    "wgt": wgt, "dx": x.grad.clone(), "grads": {k: p.grad.clone() for k, p in m.named_parameters()},
                    "act_class": type(m.act).__name__}
    torch.save(fx, os.path.join(OUT, "mlp_acts.pt"))
    print("mlp act cases", {k: (v["act_class"], float(v["y"].abs().max())) for k, v in fx.items() if isinstance(v, dict)})


MAP_KWARGS = dict(img_size=64, patch_size=16, in_chans=3, num_classes=10, embed_dim=128, depth=1, num_heads=2, mlp_ratio=2.0,
                  qkv_bias=True, class_token=False, global_pool="map")        # SigLIP-style head (vision_transformer.py:882-890)
CNX_V2_KWARGS = dict(depths=(1, 1, 2, 1), dims=(16, 32, 64, 128), num_classes=10, use_grn=True, ls_init_value=None)


def prefetch_case():
    """timm.data.loader.PrefetchLoader (run on the CPU device) over one uint8 batch: the normalised fp32 batch the model
    would receive -> tests/golden/prefetch.pt."""
    import timm
    from timm.data.loader import PrefetchLoader
    g = torch.Generator().manual_seed(71)
    batch = torch.randint(0, 256, (3, 3, 32, 32), dtype=torch.uint8, generator=g)
    target = torch.randint(0, 10, (3,), generator=g)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    (x, y), = list(PrefetchLoader([(batch, target)], mean=mean, std=std, device=torch.device("cpu")))
    fx = {"kind": "prefetch", "u8": batch, "target": target, "mean": mean, "std": std, "normalized": x.clone(),
          "timm_version": timm.__version__}
    torch.save(fx, os.path.join(OUT, "prefetch.pt"))
    print("prefetch", x.dtype, float(x.abs().max()))


# head_dim 80 (ViT-H's width: attention.py:85-93 takes dim // num_heads): the file name deliberately does not match the
# vit_*.pt glob of the GPU model tests -- the wide-head GPU tests (tests/test_gpu_wide_heads.py) compare against the oracle,
# this fixture pins the oracle at that width to the unmodified reference (tests/test_oracle_golden.py)
WIDE_KWARGS = dict(img_size=64, patch_size=16, in_chans=3, num_classes=10, embed_dim=160, depth=2, num_heads=2,
                   mlp_ratio=4.0, qkv_bias=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "wide":
        os.makedirs(OUT, exist_ok=True)
        return vit_case("wide_vit_g10_head80", 13, 2, WIDE_KWARGS)
    if len(sys.argv) > 1 and sys.argv[1] == "prefetch":
        os.makedirs(OUT, exist_ok=True)
        return prefetch_case()
    if len(sys.argv) > 1 and sys.argv[1] == "acts":
        os.makedirs(OUT, exist_ok=True)
        mlp_act_cases()
        vit_case("vit_g9_siglip_like", 11, 2, SIGLIP_KWARGS)
        return convnext_case("convnext_g3_tanh_k7553", 12, 2, CNX_TANH_KWARGS, 64, backward=True)
    if len(sys.argv) > 1 and sys.argv[1] == "map":
        os.makedirs(OUT, exist_ok=True)
        return vit_case("vit_g8_map_pool", 10, 3, MAP_KWARGS)
    if len(sys.argv) > 1 and sys.argv[1] == "convnextv2":
        os.makedirs(OUT, exist_ok=True)
        return convnext_case("convnext_g2_v2_grn", 9, 2, CNX_V2_KWARGS, 64, backward=True)
    if len(sys.argv) > 1 and sys.argv[1] == "dynamic":
        os.makedirs(OUT, exist_ok=True)
        return vit_case("vit_g7_dynamic_img_size", 8, 2, DYN_KWARGS, input_size=(96, 80))
    if len(sys.argv) > 1 and sys.argv[1] == "qknorm":
        os.makedirs(OUT, exist_ok=True)
        return qknorm_cases()
    if len(sys.argv) > 1 and sys.argv[1] == "glu":
        os.makedirs(OUT, exist_ok=True)
        return glu_cases()
    if len(sys.argv) > 1 and sys.argv[1] == "loss":
        os.makedirs(OUT, exist_ok=True)
        return loss_cases()
    if len(sys.argv) > 1 and sys.argv[1] == "optim":
        os.makedirs(OUT, exist_ok=True)
        return optim_cases()
    if len(sys.argv) > 1 and sys.argv[1] == "ema":
        os.makedirs(OUT, exist_ok=True)
        return ema_case()
    if len(sys.argv) > 1 and sys.argv[1] == "rope":
        os.makedirs(OUT, exist_ok=True)
        return rope_cases()
    if len(sys.argv) > 1 and sys.argv[1] == "naflex":
        os.makedirs(OUT, exist_ok=True)
        return naflex_case("naflex_g1", 4, 5)
    if len(sys.argv) > 1 and sys.argv[1] == "rmsnorm":
        os.makedirs(OUT, exist_ok=True)
        return vit_case("vit_g4_rmsnorm", 5, 2, RMS_KWARGS)
    os.makedirs(OUT, exist_ok=True)
    base = dict(img_size=64, patch_size=16, in_chans=3, num_classes=10, embed_dim=128, depth=2, num_heads=2,
                mlp_ratio=4.0, qkv_bias=True)
    vit_case("vit_g1_default", 0, 3, base)
    vit_case("vit_g2_layerscale_reg_avg", 1, 2,
             dict(img_size=48, patch_size=8, in_chans=3, num_classes=7, embed_dim=128, depth=1, num_heads=2,
                  qkv_bias=True, init_values=0.1, reg_tokens=2, global_pool="avg"), backward=False)
    vit_case("vit_g3_prenorm_noembedclass", 2, 2,
             dict(img_size=56, patch_size=14, in_chans=3, num_classes=5, embed_dim=64, depth=1, num_heads=1,
                  qkv_bias=True, pre_norm=True, no_embed_class=True), backward=False)
    convnext_case("convnext_g1", 3, 2, dict(depths=(1, 1, 2, 1), dims=(16, 32, 64, 128), num_classes=10), 64)
    naflex_case("naflex_g1", 4, 5)
    layer_cases()
    rope_cases()
    vit_case("vit_g4_rmsnorm", 5, 2, RMS_KWARGS)
    optim_cases()
    ema_case()
    loss_cases()
    glu_cases()
    qknorm_cases()
    vit_case("vit_g7_dynamic_img_size", 8, 2, DYN_KWARGS, input_size=(96, 80))
    convnext_case("convnext_g2_v2_grn", 9, 2, CNX_V2_KWARGS, 64, backward=True)
    vit_case("vit_g8_map_pool", 10, 3, MAP_KWARGS)
    mlp_act_cases()
    vit_case("vit_g9_siglip_like", 11, 2, SIGLIP_KWARGS)
    convnext_case("convnext_g3_tanh_k7553", 12, 2, CNX_TANH_KWARGS, 64, backward=True)
    prefetch_case()
    vit_case("wide_vit_g10_head80", 13, 2, WIDE_KWARGS)


if __name__ == "__main__":
    main()
