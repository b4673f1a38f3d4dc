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
