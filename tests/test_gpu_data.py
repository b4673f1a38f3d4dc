"""Input-pipeline tail: uint8 loader batches normalised inside the patch gather.  The reference
pipeline is PrefetchLoader's `x.to(float32).sub_(mean * 255).div_(std * 255)` (timm/data/loader.py:137-142) followed by
the model's bf16 cast; the fused gather must reproduce it BIT-EXACTLY (same fp32 operations, one rounding to bf16)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def _reference_normalise(u8):
    """What timm's PrefetchLoader does on the device (loader.py:106-109, 139), restated with torch ops."""
    mean = torch.tensor([x * 255 for x in MEAN], device=u8.device, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor([x * 255 for x in STD], device=u8.device, dtype=torch.float32).view(1, 3, 1, 1)
    return u8.to(torch.float32).sub_(mean).div_(std)


@pytest.mark.parametrize("p,size", [(16, 64), (14, 56), (4, 32), (8, 40)])
def test_patchify_u8_bit_exact_vs_normalise_then_cast(p, size):
    from b200timm import ops
    torch.manual_seed(p)
    u8 = torch.randint(0, 256, (5, 3, size, size), dtype=torch.uint8, device="cuda")
    ops.set_input_norm(MEAN, STD)
    try:
        got = ops.patchify(u8, p)
    finally:
        ops.set_input_norm(None)
    want = ops.patchify(_reference_normalise(u8).to(torch.bfloat16), p)
    assert torch.equal(got, want)
    with pytest.raises(ops.B200Error):
        ops.patchify(u8, p)                 # uint8 without a normalisation is an error, not a silent cast


def test_models_take_uint8_batches_from_the_prefetch_loader():
    from b200timm import data as BD, ops
    from b200timm.convnext import ConvNeXt
    from b200timm.vision_transformer import VisionTransformer
    torch.manual_seed(0)
    batches = [(torch.randint(0, 256, (4, 3, 64, 64), dtype=torch.uint8), torch.randint(0, 10, (4,))) for _ in range(3)]
    loader = BD.PrefetchLoader(batches, mean=MEAN, std=STD)
    assert len(loader) == 3
    vit = VisionTransformer(img_size=64, patch_size=16, embed_dim=128, depth=1, num_heads=2, num_classes=10)
    vit = vit.cuda().to(torch.bfloat16).eval()
    cnx = ConvNeXt(depths=(1, 1, 1, 1), dims=(16, 32, 64, 128), num_classes=10).cuda().to(torch.bfloat16).eval()
    try:
        seen = 0
        for (img, tgt), (ref_img, ref_tgt) in zip(loader, batches):
            assert img.dtype == torch.uint8 and img.is_cuda and torch.equal(img.cpu(), ref_img)
            assert torch.equal(tgt.cpu(), ref_tgt)
            x = _reference_normalise(img).to(torch.bfloat16)
            with torch.no_grad():
                # the patch values are bit-identical (test above); the bf16 image takes the im2col-free projection (K runs in
                # the tensor map's order), the uint8 image the gather + GEMM: fp32 sums in a different order, bf16 logits
                # within a few ulps
                a, b = vit(img).float(), vit(x).float()
                assert (a - b).abs().max().item() <= 4e-3 + 1.6e-2 * b.abs().max().item()
                assert torch.equal(cnx(img), cnx(x))
            seen += 1
        assert seen == 3
        # training through the uint8 path: same parameter gradients as through the normalised bf16 image
        vit.train()
        img = batches[0][0].cuda()
        vit(img).float().mean().backward()
        g1 = {n: p.grad.clone() for n, p in vit.named_parameters()}
        vit.zero_grad(set_to_none=True)
        vit(_reference_normalise(img).to(torch.bfloat16)).float().mean().backward()
        for n, p in vit.named_parameters():
            assert (g1[n].float() - p.grad.float()).norm() <= 1e-2 * p.grad.float().norm() + 1e-7, n
        assert torch.equal(g1["patch_embed.proj.weight"] != 0, vit.patch_embed.proj.weight.grad != 0)
    finally:
        ops.set_input_norm(None)


# ------------------------------------------------------------------------------------------------ NaFlex input tail
def _timm_batch_patchify():
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from oracle import install_ref
    if not install_ref.add_to_path():
        pytest.skip("oracle/_ref (the installed reference) is missing")
    from timm.models.naflexvit import batch_patchify
    return batch_patchify


@pytest.mark.parametrize("p,H,W", [(16, 64, 48), (14, 56, 70), (16, 50, 37), (8, 40, 40)])
@pytest.mark.parametrize("channels_last", [True, False])
def test_batch_patchify_matches_timms_function(p, H, W, channels_last):
    """b200_batch_patchify == timm.models.naflexvit.batch_patchify (the unmodified function, pad=True) on bf16 / fp32 images
    (values copied: bit-exact) and, for uint8 images, == normalise-in-fp32 -> batch_patchify -> bf16 (bit-exact too; the
    zero PADDING pixels of a ragged edge are normalised like any other zero pixel, as in the reference order of operations
    pad(normalise(x)) only when the image is patchified BEFORE normalisation -- the loader's order)."""
    from b200timm import ops
    ref_fn = _timm_batch_patchify()
    torch.manual_seed(p + H)
    img = torch.randn(3, 3, H, W, device="cuda")
    for t in (img, img.bfloat16()):
        got, grid = ops.batch_patchify(t, p, channels_last=channels_last)
        want, rgrid = ref_fn(t.float(), (p, p), pad=True, channels_last=channels_last)
        K = 3 * p * p
        assert tuple(grid) == tuple(rgrid)
        assert torch.equal(got[:, :K].float(), want.reshape(-1, K).bfloat16().float())
        assert not got[:, K:].any()
    u8 = torch.randint(0, 256, (3, 3, H, W), dtype=torch.uint8, device="cuda")
    ops.set_input_norm(MEAN, STD)
    try:
        got, _ = ops.batch_patchify(u8, p, channels_last=channels_last)
        # the NaFlex pipeline patchifies uint8 images in the dataset and normalises the PATCHES in the loader
        raw, _ = ref_fn(u8.float(), (p, p), pad=True, channels_last=channels_last)           # [B, N, K] of 0..255 values
        K = 3 * p * p
        v = raw.view(3, -1, p * p, 3) if channels_last else raw.view(3, -1, 3, p * p)
        shape = (1, 1, 1, 3) if channels_last else (1, 1, 3, 1)
        mean = torch.tensor([m * 255 for m in MEAN], device="cuda").view(shape)
        std = torch.tensor([s * 255 for s in STD], device="cuda").view(shape)
        want = v.sub(mean).div(std).reshape(-1, K).bfloat16()
        assert torch.equal(got[:, :K], want)
        # the loader's already-cut uint8 patches through b200_patches_u8: the same values
        if channels_last:
            got2 = ops.patches_u8(raw.to(torch.uint8).contiguous(), 3, channels_last=True)
            assert torch.equal(got2[:, :K], want) and not got2[:, K:].any()
    finally:
        ops.set_input_norm(None)


def test_naflex_model_takes_images_and_uint8_patch_dicts():
    """The mirror NaFlexVit on (a) an NCHW image batch (batch_patchify fused into the gather), (b) the loader's uint8 patch
    dict (normalisation fused) -- both equal the forward on the explicitly prepared bf16 patch dict."""
    from b200timm import data as BD, ops
    from b200timm.naflexvit import NaFlexVit
    ref_fn = _timm_batch_patchify()
    torch.manual_seed(0)
    model = NaFlexVit(patch_size=16, embed_dim=128, depth=2, num_heads=2, num_classes=10, pos_embed_grid_size=(4, 4))
    model = model.cuda().to(torch.bfloat16).eval()
    img = torch.randn(3, 3, 64, 48, device="cuda").bfloat16()
    patches, (gh, gw) = ref_fn(img, (16, 16))
    ys, xs = torch.meshgrid(torch.arange(gh, device="cuda"), torch.arange(gw, device="cuda"), indexing="ij")
    coord = torch.stack([ys, xs], -1).reshape(1, -1, 2).expand(3, -1, -1).contiguous()
    valid = torch.ones(3, gh * gw, dtype=torch.bool, device="cuda")
    with torch.no_grad():
        want = model(patches, coord, valid)
        got = model(img)
        assert torch.equal(got, want)
        u8 = torch.randint(0, 256, (3, 3, 64, 48), dtype=torch.uint8, device="cuda")
        raw, _ = ref_fn(u8.float(), (16, 16))
        raw_u8 = raw.to(torch.uint8).contiguous()
        loader = BD.NaFlexPrefetchLoader([({"patches": raw_u8.cpu(), "patch_coord": coord.cpu(), "patch_valid": valid.cpu()},
                                           torch.zeros(3, dtype=torch.long))], mean=MEAN, std=STD)
        try:
            (d, tgt), = list(loader)
            assert d["patches"].dtype == torch.uint8 and d["patches"].is_cuda
            got = model(d)
            mean = torch.tensor([m * 255 for m in MEAN], device="cuda").view(1, 1, 1, 3)
            std = torch.tensor([s * 255 for s in STD], device="cuda").view(1, 1, 1, 3)
            norm = raw.view(3, -1, 256, 3).sub(mean).div(std).reshape(3, -1, 768).bfloat16()
            assert torch.equal(got, model(norm, coord, valid))
        finally:
            ops.set_input_norm(None)
