"""NaFlex mode (BASELINE.json configs[4]) on the GPU: variable-length pre-patchified batches with key-padding attention
against the reference golden (tests/golden/naflex_g1.pt) and the CPU oracle at ViT-L/14 width."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _check(got, ref, atol, rtol, what, min_cos=0.999):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert got.shape == ref.shape and torch.isfinite(got).all(), what
    err = (got - ref).abs()
    cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
    assert not (err > atol + rtol * ref.abs()).any() and cos >= min_cos, (
        f"{what}: max err {err.max().item():.4g}, ref absmax {ref.abs().max().item():.4g}, cos {cos:.6f}")


def test_naflex_matches_reference_golden():
    from b200timm.naflexvit import NaFlexVit
    fx = torch.load(os.path.join(GOLDEN, "naflex_g1.pt"))
    model = NaFlexVit(**fx["kwargs"])
    model.load_state_dict(fx["state_dict"], strict=True)
    model = model.cuda().to(torch.bfloat16).eval()
    args = (fx["patches"].cuda().bfloat16(), fx["patch_coord"].cuda(), fx["patch_valid"].cuda())
    with torch.no_grad():
        feats = model.forward_features(*args)
        logits = model.forward_head(feats)
        logits2 = model({"patches": args[0], "patch_coord": args[1], "patch_valid": args[2]})
    _check(logits, fx["logits"], 2e-2, 2e-2, "naflex logits vs timm golden")
    assert torch.equal(logits, logits2)
    for b in range(args[0].shape[0]):
        n = int(fx["patch_valid"][b].sum()) + 1
        _check(feats["patches"][b, :n], fx["features"][b, :n], 6e-2, 3e-2, f"naflex features sample {b}")


def test_naflex_vit_large_width_matches_oracle():
    """ViT-L/14 geometry (D=1024, 16 heads, 577 tokens -> multi-block online-softmax attention) at depth 2."""
    from oracle import naflex_oracle as NO
    from oracle.make_golden import naflex_batch
    from b200timm.naflexvit import create_model
    torch.manual_seed(0)
    model = create_model("vit_large_patch14_clip_336", depth=2, num_classes=16)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.05)
        model.embeds.cls_token.normal_(std=0.05)
    sd = {k: v.detach().bfloat16().float() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    patches, coord, valid = naflex_batch(g, 4, (24, 24), 14)
    patches = patches.bfloat16()
    model = model.cuda().to(torch.bfloat16).eval()
    with torch.no_grad():
        logits = model(patches.cuda(), coord.cuda(), valid.cuda())
        _, ref = NO.naflex_forward(patches.float(), coord, valid, sd, dict(depth=2, num_heads=16, eps=1e-5))
    _check(logits, ref, 3e-2, 3e-2, "ViT-L/14 NaFlex logits vs oracle")


# ------------------------------------------------------------------------------------------------ training
def _rel(g, ref):
    g, ref = g.float().cpu().flatten(), ref.float().cpu().flatten()
    denom = ref.norm().item()
    cos = torch.nn.functional.cosine_similarity(g, ref, dim=0).item() if denom > 0 else 1.0
    return (g - ref).norm().item() / max(denom, 1e-12), cos, denom


def _check_grads(model, ref_grads, rel_tol=4e-2, min_cos=0.998):
    """Same criterion as tests/test_gpu_train.py: per parameter relative L2 error <= rel_tol and cosine >= min_cos;
    gradients that are negligible next to the model's largest are compared absolutely."""
    assert sorted(n for n, _ in model.named_parameters()) == sorted(ref_grads), "parameter sets differ"
    scale = max(r.float().norm().item() for r in ref_grads.values())
    for n, p in model.named_parameters():
        assert p.grad is not None, f"{n} has no grad"
        assert p.grad.dtype == p.dtype and p.grad.shape == p.shape and torch.isfinite(p.grad).all(), n
        rel, cos, denom = _rel(p.grad, ref_grads[n])
        if denom < 1e-3 * scale:
            assert (p.grad.float().cpu() - ref_grads[n].float().cpu()).norm().item() <= 4e-2 * 1e-3 * scale + 1e-7, n
            continue
        assert rel <= rel_tol and cos >= min_cos, f"{n}: rel L2 err {rel:.4f}, cos {cos:.5f}"


def test_naflex_assemble_bwd_kernel_bit_exact_copy_and_sums():
    """dproj is a bit-exact copy of the patch rows; dprefix / dtable equal an fp32 index_add of the same rows."""
    from b200timm import ops
    torch.manual_seed(0)
    B, N, P, D = 6, 40, 2, 256
    sizes = [(5, 8), (4, 4), (5, 8), (2, 3), (4, 4), (6, 6)]
    uniq, off = {}, 0
    for s in sizes:
        if s not in uniq:
            uniq[s] = off
            off += s[0] * s[1]
    pos_off = torch.tensor([uniq[s] for s in sizes], dtype=torch.int32, device="cuda")
    pos_len = torch.tensor([h * w for h, w in sizes], dtype=torch.int32, device="cuda")
    dtok = torch.randn(B, P + N, D, device="cuda").bfloat16()
    dproj, dprefix, dtable = ops.naflex_assemble_bwd(dtok, pos_off, pos_len, off, B, N, P, D)
    assert torch.equal(dproj.view(B, N, D), dtok[:, P:])
    ref_pre = dtok[:, :P].float().sum(0)
    ref_tab = torch.zeros(off, D, device="cuda")
    for b, s in enumerate(sizes):
        n = min(N, s[0] * s[1])
        ref_tab[uniq[s]:uniq[s] + n] += dtok[b, P:P + n].float()
    torch.testing.assert_close(dprefix, ref_pre, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(dtable, ref_tab, atol=1e-5, rtol=1e-5)
    # no prefix, no table: pure copy
    dproj2, dpre2, dtab2 = ops.naflex_assemble_bwd(dtok[:, P:].contiguous(), None, None, 0, B, N, 0, D)
    assert dpre2 is None and dtab2 is None and torch.equal(dproj2.view(B, N, D), dtok[:, P:])


@pytest.mark.parametrize("master_fp32", [False, True])
def test_naflex_backward_matches_reference_golden(master_fp32):
    """Every parameter gradient of NaFlexVit (incl. pos_embed through the per-sample bicubic resize, cls token, the
    linear patch projection with its 588 -> 592 padded rows) against the unmodified reference's backward."""
    from b200timm.naflexvit import NaFlexVit
    fx = torch.load(os.path.join(GOLDEN, "naflex_g1.pt"))
    model = NaFlexVit(**fx["kwargs"])
    model.load_state_dict(fx["state_dict"], strict=True)
    model = model.cuda().train()
    patches = fx["patches"].cuda()
    if not master_fp32:
        model, patches = model.to(torch.bfloat16), patches.bfloat16()
    out = model(patches, fx["patch_coord"].cuda(), fx["patch_valid"].cuda())
    assert out.requires_grad
    _check(out, fx["logits"], 2e-2, 2e-2, "naflex training-mode logits vs timm golden")
    out.float().mean().backward()
    _check_grads(model, fx["grads"])


def test_naflex_avg_pool_training_matches_oracle_autograd():
    """Masked average pooling + fc_norm head (global_pool_naflex, naflexvit.py:1077-1082), reg tokens, no pre-norm:
    loss and gradients against autograd of the fp32 oracle on the same bf16-rounded weights."""
    from oracle import naflex_oracle as NO
    from oracle import vit_oracle as O
    from oracle.make_golden import naflex_batch
    from b200timm.naflexvit import NaFlexVit
    torch.manual_seed(3)
    kw = dict(patch_size=16, embed_dim=128, depth=2, num_heads=2, class_token=True, reg_tokens=2,
              pos_embed_grid_size=(6, 6), pre_norm=False, global_pool="avg", num_classes=8)
    model = NaFlexVit(**kw)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.05)
        model.embeds.cls_token.normal_(std=0.05)
        model.embeds.reg_token.normal_(std=0.05)
    model = model.cuda().to(torch.bfloat16).train()
    g = torch.Generator().manual_seed(4)
    patches, coord, valid = naflex_batch(g, 5, (6, 6), 16)
    patches = patches.bfloat16()
    out = model(patches.cuda(), coord.cuda(), valid.cuda())
    wgt = torch.randn(out.shape, generator=g)
    (out.float() * wgt.cuda()).sum().backward()

    sd = {k: v.detach().float().cpu().requires_grad_(True) for k, v in model.state_dict().items()}
    x, npre = NO.naflex_embeds(patches.float(), coord, sd)
    kv_len = valid.sum(1) + npre
    for i in range(kw["depth"]):
        x = O.block(x, sd, f"blocks.{i}.", kw["num_heads"], 1e-6, kv_len)
    m = valid.unsqueeze(-1).float()
    pooled = (x[:, npre:] * m).sum(1) / m.sum(1).clamp(min=1)
    pooled = O.layer_norm(pooled, sd["fc_norm.weight"], sd["fc_norm.bias"], 1e-6)
    ref = O.linear(pooled, sd["head.weight"], sd["head.bias"])
    _check(out, ref.detach(), 3e-2, 3e-2, "avg-pool NaFlex logits vs oracle")
    (ref * wgt).sum().backward()
    _check_grads(model, {k: v.grad for k, v in sd.items()})
