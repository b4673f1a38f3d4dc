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
                assert torch.equal(vit(img), vit(x))
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
