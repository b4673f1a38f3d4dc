"""ConvNeXt (BASELINE.json configs[2]) on the GPU: channels-last kernels + fused GEMMs against the reference golden
(tests/golden/convnext_g1.pt, made from unmodified timm) and the CPU oracle.  Tolerances as in test_gpu_models.py."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _check(got, ref, atol, rtol, what, min_cos=0.999):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert got.shape == ref.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(got).all(), f"{what}: non-finite"
    err = (got - ref).abs()
    bad = err > atol + rtol * ref.abs()
    cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
    assert not bad.any() and cos >= min_cos, (f"{what}: {int(bad.sum())}/{bad.numel()} off, max err {err.max().item():.4g}, "
                                              f"ref absmax {ref.abs().max().item():.4g}, cos {cos:.6f}")


@pytest.mark.parametrize("k", [7, 5, 3])
@pytest.mark.parametrize("B,H,W,C", [(2, 14, 14, 512), (1, 56, 56, 128), (3, 7, 7, 1024), (2, 9, 13, 40)])
def test_dwconv_nhwc(k, B, H, W, C):
    from b200timm import ops
    g = torch.Generator(device="cuda").manual_seed(k * 100 + C)
    x = torch.randn(B, H, W, C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(C, 1, k, k, device="cuda", generator=g) * 0.2).bfloat16()
    b = torch.randn(C, device="cuda", generator=g).bfloat16()
    y = ops.dwconv_nhwc(x, ops.dwconv_weight_t(w), b)
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b.float(), padding=k // 2, groups=C)
    _check(y, ref.permute(0, 2, 3, 1), 2e-2, 1e-2, f"dwconv k={k} {B}x{H}x{W}x{C}")


def test_patchify_and_avgpool_nhwc():
    from b200timm import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(3, 8, 12, 64, device="cuda", generator=g).bfloat16()
    out = ops.patchify_nhwc(x, 2)
    ref = x.view(3, 4, 2, 6, 2, 64).permute(0, 1, 3, 2, 4, 5).reshape(3 * 4 * 6, 2 * 2 * 64)
    assert torch.equal(out, ref)
    pooled = ops.avgpool_nhwc(x)
    _check(pooled, x.float().mean((1, 2)), 1e-2, 1e-2, "avgpool")


def test_convnext_matches_reference_golden():
    from oracle import convnext_oracle as C
    from b200timm.convnext import ConvNeXt
    fx = torch.load(os.path.join(GOLDEN, "convnext_g1.pt"))
    model = ConvNeXt(**fx["kwargs"])
    model.load_state_dict(fx["state_dict"], strict=True)
    model = model.cuda().to(torch.bfloat16).eval()
    img = fx["img"].cuda().bfloat16()
    with torch.no_grad():
        feats = model.forward_features(img)
        logits = model(img)
        logits2 = model.forward_head(feats)
    _check(feats, fx["features"], 8e-2, 3e-2, "features vs timm golden")
    _check(logits, fx["logits"], 2e-2, 2e-2, "logits vs timm golden")
    _check(logits2, fx["logits"], 2e-2, 2e-2, "forward_head(forward_features) vs timm golden")
    sd = {k: v.bfloat16().float() for k, v in fx["state_dict"].items()}
    ref = C.convnext_forward(fx["img"].bfloat16().float(), sd, dict(fx["kwargs"]))
    _check(logits, ref, 2e-2, 2e-2, "logits vs oracle(bf16 weights)")


def test_convnext_base_matches_oracle():
    """BASELINE configs[2] architecture (convnext_base @224) at batch 4 against the fp32 oracle, with the
    SURVEY 8c yardstick |ours - fp32| <= max(2 |eager_bf16 - fp32|, atol)."""
    from oracle import convnext_oracle as C
    from b200timm.convnext import create_model
    cfg = dict(C.CONVNEXT_CONFIGS["convnext_base"])
    sd = {k: v.bfloat16().float() for k, v in C.random_state_dict(cfg, seed=2).items()}
    model = create_model("convnext_base")
    model.load_state_dict(sd, strict=True)
    model = model.cuda().to(torch.bfloat16).eval()
    img = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(3)).bfloat16()
    with torch.no_grad():
        logits = model(img.cuda()).float().cpu()
        ref = C.convnext_forward_eager(img.float(), sd, cfg)
        eager16 = C.convnext_forward_eager(img.cuda(), {k: v.cuda().bfloat16() for k, v in sd.items()}, cfg).float().cpu()
    e_ref, e_ours = (eager16 - ref).abs().max().item(), (logits - ref).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(logits.flatten(), ref.flatten(), dim=0).item()
    assert torch.isfinite(logits).all()
    assert e_ours <= max(2 * e_ref, 2e-2) and cos >= 0.9995, (
        f"convnext_base: max err {e_ours:.4g} vs eager-bf16 {e_ref:.4g} (ref absmax {ref.abs().max().item():.3g}), cos {cos:.6f}")


def test_dwconv_backward_kernels():
    from b200timm import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    B, H, W, C, k = 2, 14, 14, 64, 7
    x = torch.randn(B, H, W, C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(C, 1, k, k, device="cuda", generator=g) * 0.2).bfloat16()
    dy = torch.randn(B, H, W, C, device="cuda", generator=g).bfloat16()
    res = torch.randn(B, H, W, C, device="cuda", generator=g).bfloat16()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.float().requires_grad_(True)
    torch.nn.functional.conv2d(xr, wr, None, padding=k // 2, groups=C).backward(dy.float().permute(0, 3, 1, 2))
    wt = ops.dwconv_weight_t(w)
    dx = ops.dwconv_nhwc(dy, wt.flip(0, 1).contiguous(), None, residual=res)
    _check(dx, xr.grad.permute(0, 2, 3, 1) + res.float(), 3e-2, 1e-2, "dwconv dgrad (+ shortcut gradient)")
    dwt = ops.dwconv_nhwc_wgrad(x, dy, k)
    _check(dwt.reshape(k * k, C).t().reshape(C, 1, k, k), wr.grad, 6e-2, 1e-2, "dwconv wgrad")
    p = ops.patchify_nhwc(x, 2)
    assert torch.equal(ops.unpatchify_nhwc(p, B, H, W, C, 2), x)


def test_convnext_backward_matches_oracle_autograd():
    """Every parameter gradient of a small ConvNeXt (depthwise conv, LayerNorm, MLP, LayerScale gamma, downsample,
    stem, head) after a cross-entropy loss, against PyTorch autograd of the fp32 oracle."""
    from oracle import convnext_oracle as C
    from b200timm.convnext import ConvNeXt
    cfg = dict(depths=(1, 1, 2, 1), dims=(32, 64, 128, 256))
    sd = {k: v.bfloat16().float() for k, v in C.random_state_dict(cfg, seed=11, num_classes=16).items()}
    model = ConvNeXt(num_classes=16, **cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().to(torch.bfloat16).train()
    g = torch.Generator().manual_seed(6)
    img = torch.randn(4, 3, 64, 64, generator=g).bfloat16()
    target = torch.randint(0, 16, (4,), generator=g)
    loss = torch.nn.functional.cross_entropy(model(img.cuda()).float(), target.cuda())
    loss.backward()
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref_loss = torch.nn.functional.cross_entropy(C.convnext_forward(img.float(), ref_sd, cfg), target)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 3e-2, (loss.item(), ref_loss.item())
    scale = max(v.grad.float().norm().item() for v in ref_sd.values())
    for n, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
        ref = ref_sd[n].grad
        denom = ref.norm().item()
        if denom < 1e-3 * scale:
            continue
        rel = (p.grad.float().cpu() - ref).norm().item() / denom
        cos = torch.nn.functional.cosine_similarity(p.grad.float().cpu().flatten(), ref.flatten(), dim=0).item()
        assert rel <= 5e-2 and cos >= 0.998, f"{n}: rel L2 err {rel:.4f}, cos {cos:.5f}"


def test_convnext_droppath_training_matches_oracle():
    """Stochastic depth in TRAIN mode (convnext.py:213): the per-sample keep masks come from the torch CUDA generator
    exactly as timm/layers/drop.py:170-175 draws them (B Bernoulli draws per active block, in block order), so
    re-seeding reproduces them for the oracle; loss and every parameter gradient must match its autograd."""
    from oracle import convnext_oracle as C
    from b200timm.convnext import ConvNeXt
    cfg = dict(depths=(1, 1, 2, 1), dims=(32, 64, 128, 256))
    sd = {k: v.bfloat16().float() for k, v in C.random_state_dict(cfg, seed=12, num_classes=16).items()}
    for k in sd:
        if k.endswith("gamma"):
            sd[k] = torch.full_like(sd[k], 0.5)          # LayerScale large enough to matter
    model = ConvNeXt(num_classes=16, drop_path_rate=0.4, **cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().to(torch.bfloat16).train()
    B = 16
    g = torch.Generator().manual_seed(9)
    img = torch.randn(B, 3, 64, 64, generator=g).bfloat16()
    target = torch.randint(0, 16, (B,), generator=g)
    torch.manual_seed(4321)
    loss = torch.nn.functional.cross_entropy(model(img.cuda()).float(), target.cuda())
    loss.backward()
    torch.manual_seed(4321)
    keeps = {}
    for i, stage in enumerate(model.stages):
        for j, blk in enumerate(stage.blocks):
            r = getattr(blk.drop_path, "drop_prob", 0.0)
            if r > 0:
                m = torch.empty((B, 1, 1, 1), device="cuda", dtype=torch.bfloat16).bernoulli_(1 - r).reshape(-1).float().cpu()
                keeps[f"stages.{i}.blocks.{j}."] = (m, 1 - r)
    assert any(0 < m.sum() < B for m, _ in keeps.values()), "test should exercise dropped and kept samples"
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    feats = C.convnext_forward_features(img.float(), ref_sd, cfg, keeps=keeps)
    ref_loss = torch.nn.functional.cross_entropy(C.convnext_forward_head(feats, ref_sd, cfg), target)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 3e-2, (loss.item(), ref_loss.item())
    scale = max(v.grad.float().norm().item() for v in ref_sd.values())
    for n, p in model.named_parameters():
        ref = ref_sd[n].grad
        if ref.norm().item() < 1e-3 * scale:
            continue
        rel = (p.grad.float().cpu() - ref).norm().item() / ref.norm().item()
        assert rel <= 5e-2, f"{n}: rel L2 err {rel:.4f}"


def test_convnext_amp_fp32_master_weights():
    """train.py --amp mode on ConvNeXt: fp32 parameters + fp32 images under torch.autocast(bfloat16); fp32 gradients for
    every parameter, matching the oracle evaluated on the bf16-rounded GEMM / conv weights."""
    from oracle import convnext_oracle as C
    from b200timm.convnext import ConvNeXt
    cfg = dict(depths=(1, 1, 2, 1), dims=(32, 64, 128, 256))
    sd = C.random_state_dict(cfg, seed=14, num_classes=16)
    model = ConvNeXt(num_classes=16, **cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    g = torch.Generator().manual_seed(3)
    img = torch.randn(4, 3, 64, 64, generator=g)
    target = torch.randint(0, 16, (4,), generator=g)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = torch.nn.functional.cross_entropy(model(img.cuda()).float(), target.cuda())
    loss.backward()
    assert all(p.grad is not None and p.grad.dtype == torch.float32 for p in model.parameters())
    ref_sd = {k: ((v.bfloat16().float() if (k.endswith("weight") and v.dim() >= 2) else v.clone())).requires_grad_(True)
              for k, v in sd.items()}
    ref_loss = torch.nn.functional.cross_entropy(C.convnext_forward(img.bfloat16().float(), ref_sd, cfg), target)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 3e-2, (loss.item(), ref_loss.item())
    scale = max(v.grad.float().norm().item() for v in ref_sd.values())
    for n, p in model.named_parameters():
        ref = ref_sd[n].grad
        if ref.norm().item() < 1e-3 * scale:
            continue
        rel = (p.grad.cpu() - ref).norm().item() / ref.norm().item()
        assert rel <= 5e-2, f"{n}: rel L2 err {rel:.4f}"
    model.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        assert torch.isfinite(model(img.cuda()).float()).all()


def test_grn_kernels_match_torch():
    """b200_grn_fwd / b200_grn_bwd against the reference formula (timm/layers/grn.py:46-49) and its autograd, with GELU in
    front as in GlobalResponseNormMlp (the backward kernel returns the gradient w.r.t. the pre-activation)."""
    from b200timm import ops
    torch.manual_seed(0)
    B, HW, C = 5, 77, 136
    pre = torch.randn(B * HW, C, device="cuda").bfloat16()
    w = (torch.randn(C, device="cuda") * 0.3).bfloat16()
    b = (torch.randn(C, device="cuda") * 0.1).bfloat16()
    d = torch.randn(B * HW, C, device="cuda").bfloat16()
    pf = pre.float().requires_grad_(True)
    wf, bf = w.float().requires_grad_(True), b.float().requires_grad_(True)
    h = torch.nn.functional.gelu(pf).view(B, HW, C)
    gx = h.norm(p=2, dim=1, keepdim=True)
    nx = gx / (gx.mean(dim=-1, keepdim=True) + 1e-6)
    ref = (h + torch.addcmul(bf.view(1, 1, -1), wf.view(1, 1, -1), h * nx)).view(B * HW, C)
    ref.backward(d.float())
    act = torch.nn.functional.gelu(pre.float()).bfloat16()
    gsq, scale = ops.grn_fwd(act, w, b, B, HW)
    _check(act, ref.detach(), 3e-2, 2e-2, "grn forward")
    dpre, dw = ops.grn_bwd(d, pre, gsq, scale, w, B, HW)
    _check(dpre, pf.grad, 4e-2, 4e-2, "grn + gelu backward", min_cos=0.9995)
    _check(dw, wf.grad, 2e-2 * wf.grad.abs().max().item(), 2e-2, "grn weight gradient")


def test_convnext_v2_grn_matches_reference_golden():
    """ConvNeXt-V2 blocks (use_grn=True, no LayerScale): inference and a training step against the reference's logits
    and parameter gradients (tests/golden/convnext_g2_v2_grn.pt)."""
    from b200timm.convnext import ConvNeXt
    from test_gpu_train import _check_grads
    fx = torch.load(os.path.join(GOLDEN, "convnext_g2_v2_grn.pt"))
    model = ConvNeXt(**fx["kwargs"])
    model.load_state_dict(fx["state_dict"], strict=True)
    model = model.cuda().to(torch.bfloat16)
    img = fx["img"].cuda().bfloat16()
    with torch.no_grad():
        feats = model.eval().forward_features(img)
        logits = model(img)
    _check(feats, fx["features"], 8e-2, 3e-2, "V2 features vs timm golden")
    _check(logits, fx["logits"], 2e-2, 2e-2, "V2 logits vs timm golden")
    out = model.train()(img)
    _check(out, fx["logits"], 2e-2, 2e-2, "V2 training-mode logits vs timm golden")
    out.float().mean().backward()
    _check_grads(model, fx["grads"], rel_tol=6e-2, min_cos=0.997)



@pytest.mark.parametrize("B,H,W,C", [(2, 14, 14, 512), (1, 56, 56, 128), (3, 7, 7, 1024), (2, 30, 17, 192)])
def test_dwconv_with_layernorm_statistics(B, H, W, C):
    """b200_dwconv_nhwc_ln: the same output bits as b200_dwconv_nhwc plus, per pixel and 64-channel block, the sum and the
    sum of squares of the stored bf16 values (the producer side of the LayerNorm folded into fc1)."""
    from b200timm import ops
    g = torch.Generator(device="cuda").manual_seed(C + H)
    x = torch.randn(B, H, W, C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(C, 1, 7, 7, device="cuda", generator=g) * 0.2).bfloat16()
    b = torch.randn(C, device="cuda", generator=g).bfloat16()
    wt = ops.dwconv_weight_t(w)
    y, st = ops.dwconv_nhwc_ln(x, wt, b)
    assert torch.equal(y, ops.dwconv_nhwc(x, wt, b))
    assert tuple(st.shape) == (C // 64 + 1, B * H * W, 2)
    yf = y.float().view(B * H * W, C // 64, 64)
    _check(st[:-1, :, 0].t(), yf.sum(-1), 1e-3, 1e-5, "partial sums", min_cos=0.99999)
    _check(st[:-1, :, 1].t(), (yf * yf).sum(-1), 1e-3, 1e-5, "partial sums of squares", min_cos=0.99999)


def test_convnext_block_with_and_without_folded_layernorm():
    """ConvNeXtBlock inference with the LayerNorm folded between the depthwise convolution and fc1 (default) and with the
    LayerNorm kernel: same output within bf16 rounding, both within tolerance of fp32 PyTorch math (convnext.py:197-213)."""
    from b200timm import functional as BF, ops
    from b200timm.convnext import ConvNeXtBlock
    torch.manual_seed(3)
    for C, HW in ((128, 28), (256, 14)):
        blk = ConvNeXtBlock(C, ls_init_value=0.5).cuda().eval()
        with torch.no_grad():
            blk.norm.weight.normal_(1.0, 0.2)
            blk.norm.bias.normal_(0.0, 0.2)
            for p in blk.parameters():
                p.copy_(p.bfloat16().float())
        x = torch.randn(2, C, HW, HW, device="cuda").bfloat16()
        with torch.no_grad():
            xf = x.float()
            r = torch.nn.functional.conv2d(xf, blk.conv_dw.weight, blk.conv_dw.bias, padding=3, groups=C).permute(0, 2, 3, 1)
            r = torch.nn.functional.layer_norm(r, (C,), blk.norm.weight, blk.norm.bias, blk.norm.eps)
            r = torch.nn.functional.gelu(r @ blk.mlp.fc1.weight.t() + blk.mlp.fc1.bias) @ blk.mlp.fc2.weight.t() + blk.mlp.fc2.bias
            ref = xf + (r * blk.gamma).permute(0, 3, 1, 2)
        blk = blk.to(torch.bfloat16)
        outs = {}
        for fold in (True, False):
            BF._LN_FOLD = fold
            try:
                ops.reset_counters()
                with torch.no_grad():
                    outs[fold] = blk(x).float()
                outs[(fold, "n")] = ops.launch_count()
            finally:
                BF._LN_FOLD = True
            _check(outs[fold], ref, 3e-2, 2e-2, f"ConvNeXtBlock C={C} fold={fold}")
        assert outs[(True, "n")] == outs[(False, "n")], outs      # LayerNorm launch <-> statistics finalisation
        _check(outs[True], outs[False], 2e-2, 1e-2, "folded vs LayerNorm kernel")
