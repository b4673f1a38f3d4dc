"""Classification losses through the C ABI: b200timm.loss.{CrossEntropyLoss,
LabelSmoothingCrossEntropy, SoftTargetCrossEntropy} against the reference values / gradients of tests/golden/loss.pt
(nn.CrossEntropyLoss and timm.loss classes, fp32) and, for bf16 logits, against the oracle on the same rounded logits.
Tolerances: fp32 logits -- loss rtol 2e-6, gradient atol 1e-8 + rtol 1e-5; bf16 logits -- loss computed in fp32 from the
bf16 values (rtol 2e-6 vs the oracle on those values), gradient rounded to bf16 (one ulp: rtol 2^-7)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _loss_modules():
    from b200timm import loss as BL
    return {"ce": (BL.CrossEntropyLoss(), "target", 0.0), "ls": (BL.LabelSmoothingCrossEntropy(0.1), "target", 0.1),
            "soft": (BL.SoftTargetCrossEntropy(), "soft_target", 0.0)}


@pytest.mark.parametrize("name", ["small", "imagenet", "wide"])
def test_fp32_logits_match_reference(name):
    case = torch.load(os.path.join(GOLDEN, "loss.pt"))["cases"][name]
    for key, (mod, tkey, _) in _loss_modules().items():
        x = case["logits"].cuda().requires_grad_(True)
        loss = mod(x, case[tkey].cuda())
        assert loss.dtype == torch.float32 and loss.dim() == 0
        (loss * case[key]["grad_out"]).backward()
        torch.testing.assert_close(loss.detach().cpu(), case[key]["loss"], atol=1e-6, rtol=2e-6, msg=lambda m: f"{key}: {m}")
        torch.testing.assert_close(x.grad.cpu(), case[key]["dlogits"], atol=1e-8, rtol=1e-5, msg=lambda m: f"{key}: {m}")


@pytest.mark.parametrize("name", ["small", "imagenet"])
def test_bf16_logits_match_oracle(name):
    from oracle import optim_oracle as OO
    case = torch.load(os.path.join(GOLDEN, "loss.pt"))["cases"][name]
    for key, (mod, tkey, s) in _loss_modules().items():
        xb = case["logits"].bfloat16()
        tgt = case[tkey]
        if tkey == "soft_target":
            tgt = tgt.bfloat16()
        x = xb.cuda().requires_grad_(True)
        loss = mod(x, tgt.cuda())
        loss.backward()
        xr = xb.float().requires_grad_(True)
        ref = OO.cross_entropy(xr, tgt, s)
        ref.backward()
        torch.testing.assert_close(loss.detach().cpu(), ref.detach(), atol=1e-6, rtol=2e-6)
        assert x.grad.dtype == torch.bfloat16
        torch.testing.assert_close(x.grad.float().cpu(), xr.grad, atol=1e-9, rtol=2.0 ** -7)


def test_training_step_with_kernel_loss_matches_torch_loss():
    from b200timm import loss as BL
    from b200timm.vision_transformer import VisionTransformer
    kw = dict(img_size=64, patch_size=16, embed_dim=128, depth=1, num_heads=2, num_classes=10)
    torch.manual_seed(0)
    m = VisionTransformer(**kw).cuda().to(torch.bfloat16).train()
    x = torch.randn(6, 3, 64, 64, device="cuda").bfloat16()
    y = torch.randint(0, 10, (6,), device="cuda")
    BL.CrossEntropyLoss()(m(x), y).backward()
    g1 = {n: p.grad.clone() for n, p in m.named_parameters()}
    m.zero_grad(set_to_none=True)
    torch.nn.functional.cross_entropy(m(x).float(), y).backward()
    for n, p in m.named_parameters():
        a, b = g1[n].float(), p.grad.float()
        assert (a - b).norm() <= 2e-2 * b.norm() + 1e-6, n


def test_loss_rejects_cpu_and_bad_targets():
    from b200timm import loss as BL
    from b200timm._lib import B200Error
    with pytest.raises(B200Error):
        BL.CrossEntropyLoss()(torch.randn(2, 5), torch.tensor([0, 1]))
    with pytest.raises(B200Error):
        BL.CrossEntropyLoss()(torch.randn(2, 5, device="cuda"), torch.tensor([0, 1], device="cuda", dtype=torch.int32))
    with pytest.raises(B200Error):
        BL.CrossEntropyLoss(reduction="sum")
