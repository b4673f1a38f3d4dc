"""MLP activations other than the exact GELU on the GEMM epilogues (b200_gemm_desc.act): tanh GELU (timm 'gelu_tanh':
SigLIP ViTs, timm's test_convnext family) and SiLU, forward (EPI_BIAS_GELU) and backward (EPI_DGELU), against the
reference goldens (tests/golden/mlp_acts.pt, vit_g9_siglip_like.pt, convnext_g3_tanh_k7553.pt)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(g, ref):
    g, ref = g.float().cpu().flatten(), ref.float().cpu().flatten()
    return (g - ref).norm().item() / max(ref.norm().item(), 1e-12), torch.nn.functional.cosine_similarity(g, ref, dim=0).item()


@pytest.mark.parametrize("act", ["gelu", "gelu_tanh", "silu"])
@pytest.mark.parametrize("M", [50432 // 16, 300])
def test_gemm_activation_epilogues_match_torch(act, M):
    """Both GEMM shapes the dispatch distinguishes (pair kernels for M >= 256 with a wide N, single-CTA otherwise)."""
    from b200timm import ops
    torch.manual_seed(0)
    K, N = 256, 512
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.08).bfloat16()
    b = (torch.randn(N, device="cuda") * 0.5).bfloat16()
    f = {"gelu": torch.nn.functional.gelu, "gelu_tanh": lambda t: torch.nn.functional.gelu(t, approximate="tanh"),
         "silu": torch.nn.functional.silu}[act]
    pre_ref = (x.float() @ w.float().t() + b.float())
    pre = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    y = ops.linear_fwd(x, w, b, epilogue=ops.EPI_BIAS_GELU, out2=pre, act=ops.ACTS[act])
    torch.testing.assert_close(pre.float(), pre_ref, atol=2e-2, rtol=2e-2)
    torch.testing.assert_close(y.float(), f(pre_ref), atol=2e-2, rtol=2e-2)
    # backward epilogue: (dY W) * act'(pre)
    dy = torch.randn(M, K, device="cuda").bfloat16()
    w2 = (torch.randn(K, N, device="cuda") * 0.08).bfloat16()          # fc2.weight [out=K, in=N]
    pf = pre.float().requires_grad_(True)
    f(pf).backward(dy.float() @ w2.float())
    got = ops.linear_dgrad(dy, w2, epilogue=ops.EPI_DGELU, aux=pre, act=ops.ACTS[act])
    rel, cos = _rel(got, pf.grad)
    assert rel <= 1.5e-2 and cos >= 0.9998, (act, rel, cos)


@pytest.mark.parametrize("name", ["gelu_tanh", "silu"])
def test_mlp_activation_modules_match_reference_golden(name):
    from b200timm import layers as L
    from functools import partial
    fx = torch.load(os.path.join(GOLDEN, "mlp_acts.pt"))[name]
    act = partial(torch.nn.GELU, approximate="tanh") if name == "gelu_tanh" else torch.nn.SiLU
    m = L.Mlp(64, 256, act_layer=act)
    m.load_state_dict(fx["state_dict"], strict=True)
    m = m.cuda().to(torch.bfloat16)
    x = fx["x"].cuda().bfloat16()
    with torch.no_grad():
        rel, cos = _rel(m.eval()(x), fx["y"])
    assert rel <= 2e-2 and cos >= 0.9995, (rel, cos)
    x.requires_grad_(True)
    y = m.train()(x)
    (y.float() * fx["wgt"].cuda()).sum().backward()
    rel, cos = _rel(x.grad, fx["dx"])
    assert rel <= 4e-2 and cos >= 0.998, ("dx", rel, cos)
    for k, p in m.named_parameters():
        rel, cos = _rel(p.grad, fx["grads"][k])
        assert rel <= 4e-2 and cos >= 0.998, (k, rel, cos)


def test_siglip_like_vit_matches_reference_golden():
    """tanh-GELU blocks + attention-pool head without a class token (the SigLIP ViT layout)."""
    from b200timm.vision_transformer import VisionTransformer
    from test_gpu_train import _check_grads
    fx = torch.load(os.path.join(GOLDEN, "vit_g9_siglip_like.pt"))
    model = VisionTransformer(**fx["kwargs"])
    model.load_state_dict(fx["state_dict"], strict=True)
    model = model.cuda().to(torch.bfloat16)
    img = fx["img"].cuda().bfloat16()
    with torch.no_grad():
        out = model.eval()(img)
    torch.testing.assert_close(out.float().cpu(), fx["logits"], atol=6e-3, rtol=3e-2)
    out = model.train()(img)
    out.float().mean().backward()
    _check_grads(model, fx["grads"])


def test_convnext_tanh_gelu_mixed_kernel_sizes_matches_reference_golden():
    """timm's test_convnext family: act_layer='gelu_tanh', depthwise kernels 7 / 5 / 5 / 3."""
    from b200timm.convnext import ConvNeXt
    from test_gpu_train import _check_grads
    fx = torch.load(os.path.join(GOLDEN, "convnext_g3_tanh_k7553.pt"))
    model = ConvNeXt(**fx["kwargs"])
    model.load_state_dict(fx["state_dict"], strict=True)
    model = model.cuda().to(torch.bfloat16)
    img = fx["img"].cuda().bfloat16()
    with torch.no_grad():
        logits = model.eval()(img)
    torch.testing.assert_close(logits.float().cpu(), fx["logits"], atol=2e-2, rtol=2e-2)
    out = model.train()(img)
    out.float().mean().backward()
    _check_grads(model, fx["grads"], rel_tol=6e-2, min_cos=0.997)
