"""Attention with qk_norm / scale_norm (row a4, timm/layers/attention.py:101-103,117,136) on the GPU: the in-place q / k
normalisation kernels and the module / model paths against the reference goldens (tests/golden/qknorm.pt,
vit_g6_qknorm.pt).  bf16 tolerances as in tests/test_gpu_train.py."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(g, ref):
    g, ref = g.float().cpu().flatten(), ref.float().cpu().flatten()
    return (g - ref).norm().item() / max(ref.norm().item(), 1e-12), torch.nn.functional.cosine_similarity(g, ref, dim=0).item()


@pytest.mark.parametrize("rms", [False, True])
@pytest.mark.parametrize("T,H", [(197 * 3, 12), (50, 2), (7, 3)])
def test_qk_norm_kernels_match_torch(rms, T, H):
    from b200timm import ops
    torch.manual_seed(0)
    qkv = torch.randn(T, 3, H, 64, device="cuda").bfloat16()
    qw, kw = (torch.randn(64, device="cuda") * 0.2 + 1).bfloat16(), (torch.randn(64, device="cuda") * 0.2 + 1).bfloat16()
    qb = kb = None
    if not rms:
        qb, kb = (torch.randn(64, device="cuda") * 0.1).bfloat16(), (torch.randn(64, device="cuda") * 0.1).bfloat16()
    eps = ops.RmsEps(1e-6) if rms else 1e-6

    def ref_norm(x, w, b):
        if rms:
            return x * torch.rsqrt((x * x).mean(-1, keepdim=True) + 1e-6) * w.float()
        return torch.nn.functional.layer_norm(x, (64,), w.float(), b.float(), 1e-6)

    x = qkv.float().requires_grad_(True)
    ref = torch.stack([ref_norm(x[:, 0], qw, qb), ref_norm(x[:, 1], kw, kb), x[:, 2]], 1)
    buf = qkv.clone().view(T, -1)
    raw, mean, rstd = ops.qk_norm_fwd(buf, T, H, qw, qb, kw, kb, eps, save=True)
    got = buf.view(T, 3, H, 64)
    assert torch.equal(got[:, 2], qkv[:, 2]) and torch.equal(raw, qkv[:, :2])
    torch.testing.assert_close(got.float(), ref.detach(), atol=3e-2, rtol=2e-2)
    buf2 = qkv.clone().view(T, -1)
    assert ops.qk_norm_fwd(buf2, T, H, qw, qb, kw, kb, eps) == (None, None, None) and torch.equal(buf2, buf)
    d = torch.randn(T, 3, H, 64, device="cuda").bfloat16()
    qwf, kwf = qw.float().requires_grad_(True), kw.float().requires_grad_(True)
    x2 = qkv.float().requires_grad_(True)
    qbf = None if rms else qb.float().requires_grad_(True)
    kbf = None if rms else kb.float().requires_grad_(True)
    r2 = torch.stack([ref_norm(x2[:, 0], qwf, qbf) if not rms else x2[:, 0] * torch.rsqrt((x2[:, 0] ** 2).mean(-1, keepdim=True) + 1e-6) * qwf,
                      ref_norm(x2[:, 1], kwf, kbf) if not rms else x2[:, 1] * torch.rsqrt((x2[:, 1] ** 2).mean(-1, keepdim=True) + 1e-6) * kwf,
                      x2[:, 2]], 1)
    r2.backward(d.float())
    dbuf = d.clone().view(T, -1)
    dp = ops.qk_norm_bwd(dbuf, raw, mean, rstd, qw, kw, T, H)
    rel, cos = _rel(dbuf.view(T, 3, H, 64), x2.grad)
    assert rel <= 2e-2 and cos >= 0.9995, (rel, cos)
    for got_p, ref_p in ((dp[0], qwf.grad), (dp[2], kwf.grad)) + (() if rms else ((dp[1], qbf.grad), (dp[3], kbf.grad))):
        rel, cos = _rel(got_p, ref_p)
        assert rel <= 2e-2 and cos >= 0.9995, (rel, cos)


def _attention(name, fx):
    from b200timm import layers as L
    kw = {"ln_qk": dict(qkv_bias=True, qk_norm=True, norm_layer=L.LayerNorm),
          "rms_qk_scale": dict(qkv_bias=False, qk_norm=True, scale_norm=True, norm_layer=L.RmsNorm),
          "ln_scale_only": dict(qkv_bias=True, scale_norm=True, norm_layer=L.LayerNorm)}[name]
    m = L.Attention(128, num_heads=2, **kw)
    m.load_state_dict(fx["state_dict"], strict=True)
    return m.cuda().to(torch.bfloat16)


@pytest.mark.parametrize("name", ["ln_qk", "rms_qk_scale", "ln_scale_only"])
@pytest.mark.parametrize("train", [False, True])
def test_attention_norm_modules_match_reference_golden(name, train):
    fx = torch.load(os.path.join(GOLDEN, "qknorm.pt"))[name]
    m = _attention(name, fx)
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
    scale = max(g.norm().item() for g in fx["grads"].values())
    for k, p in m.named_parameters():
        ref = fx["grads"][k]
        assert p.grad.dtype == p.dtype
        if ref.norm().item() < 1e-3 * scale:
            # analytically zero: a bias on k shifts every score of a query by the same q.b, which softmax ignores; the
            # reference holds fp32 cancellation noise there, the kernels bf16 cancellation noise
            assert k == "k_norm.bias" and p.grad.float().norm().item() <= 2e-2 * scale, (k, p.grad.float().norm().item(), scale)
            continue
        rel, cos = _rel(p.grad, ref)
        assert rel <= 4e-2 and cos >= 0.998, (k, rel, cos)


@pytest.mark.parametrize("train", [False, True])
def test_qknorm_vit_matches_reference_golden(train):
    from b200timm.vision_transformer import VisionTransformer
    from test_gpu_train import _check_grads
    fx = torch.load(os.path.join(GOLDEN, "vit_g6_qknorm.pt"))
    model = VisionTransformer(**fx["kwargs"])
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
