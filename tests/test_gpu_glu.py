"""Gated MLPs (SwiGLU / SwiGLUPacked / GluMlp, timm/layers/mlp.py:57-165) on the GPU: the kernels'
forward, input gradient and parameter gradients against the reference goldens (tests/golden/glu.pt,
vit_g5_swiglu.pt).  bf16 tolerances as in tests/test_gpu_train.py: relative L2 error <= 4e-2, cosine >= 0.998."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(g, ref):
    g, ref = g.float().cpu().flatten(), ref.float().cpu().flatten()
    return (g - ref).norm().item() / max(ref.norm().item(), 1e-12), torch.nn.functional.cosine_similarity(g, ref, dim=0).item()


def _module(name):
    from b200timm import layers as L
    return {"swiglu": lambda: L.SwiGLU(64, 128), "swiglu_norm": lambda: L.SwiGLU(64, 128, norm_layer=L.LayerNorm),
            "swiglu_packed": lambda: L.SwiGLUPacked(64, 256),
            "glu_sigmoid": lambda: L.GluMlp(64, 256, bias=(True, False))}[name]()


def test_glu_kernels_match_torch_elementwise():
    from b200timm import ops
    torch.manual_seed(0)
    h = torch.randn(300, 2 * 136, device="cuda").bfloat16()
    d = torch.randn(300, 136, device="cuda").bfloat16()
    for gate_first in (True, False):
        for act, f in (("silu", torch.nn.functional.silu), ("sigmoid", torch.sigmoid)):
            hf = h.float().requires_grad_(True)
            a, b = hf.chunk(2, -1)
            ref = f(a) * b if gate_first else a * f(b)
            got = ops.glu_fwd(h, gate_first, act)
            torch.testing.assert_close(got.float(), ref.detach(), atol=1e-2, rtol=1e-2)
            assert (got.float() - ref.detach()).abs().max() <= 2.0 ** -7 * ref.abs().max()
            ref.backward(d.float())
            dh = ops.glu_bwd(h, d, gate_first, act)
            torch.testing.assert_close(dh.float(), hf.grad, atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("name", ["swiglu", "swiglu_norm", "swiglu_packed", "glu_sigmoid"])
@pytest.mark.parametrize("train", [False, True])
def test_glu_modules_match_reference_golden(name, train):
    fx = torch.load(os.path.join(GOLDEN, "glu.pt"))[name]
    m = _module(name)
    m.load_state_dict(fx["state_dict"], strict=True)
    m = m.cuda().to(torch.bfloat16)
    x = fx["x"].cuda().bfloat16()
    if not train:
        with torch.no_grad():
            y = m.eval()(x)
        rel, cos = _rel(y, fx["y"])
        assert rel <= 2e-2 and cos >= 0.9995, (rel, cos)
        return
    x.requires_grad_(True)
    y = m.train()(x)
    (y.float() * fx["wgt"].cuda()).sum().backward()
    rel, cos = _rel(y, fx["y"])
    assert rel <= 2e-2 and cos >= 0.9995, (rel, cos)
    rel, cos = _rel(x.grad, fx["dx"])
    assert rel <= 4e-2 and cos >= 0.998, ("dx", rel, cos)
    for k, p in m.named_parameters():
        rel, cos = _rel(p.grad, fx["grads"][k])
        assert p.grad.dtype == p.dtype and rel <= 4e-2 and cos >= 0.998, (k, rel, cos)


@pytest.mark.parametrize("train", [False, True])
def test_swiglu_vit_matches_reference_golden(train):
    """VisionTransformer(mlp_layer=SwiGLU, act_layer=SiLU, LayerScale): inference through the fused block (residual in
    fc2's epilogue), training through the composed per-module autograd functions."""
    from b200timm import layers as L
    from b200timm.vision_transformer import VisionTransformer
    from test_gpu_train import _check_grads
    fx = torch.load(os.path.join(GOLDEN, "vit_g5_swiglu.pt"))
    kw = dict(fx["kwargs"])          # mlp_layer="SwiGLU", act_layer="silu": resolved by the mirror
    model = VisionTransformer(**kw)
    model.load_state_dict(fx["state_dict"], strict=True)
    model = model.cuda().to(torch.bfloat16)
    img = fx["img"].cuda().bfloat16()
    if not train:
        with torch.no_grad():
            out = model.eval()(img)
        torch.testing.assert_close(out.float().cpu(), fx["logits"], atol=2e-2, rtol=2e-2)
        return
    out = model.train()(img)
    torch.testing.assert_close(out.float().cpu(), fx["logits"], atol=2e-2, rtol=2e-2)
    out.float().mean().backward()
    _check_grads(model, fx["grads"])
