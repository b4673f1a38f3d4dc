"""Rotary-embedding attention: timm.layers.AttentionRope (attention.py:224-282) through the fused
QKV + RoPE GEMM epilogue, the in-place rotation kernel and the autograd path, against the oracle / golden fixture."""
import os

import pytest
import torch

from oracle import vit_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _emb(n_rot, seed=0):
    g = torch.Generator().manual_seed(seed)
    ang = torch.rand(n_rot, 32, generator=g) * 6.0
    ang = ang.repeat_interleave(2, dim=-1)                      # interleaved tables duplicate each frequency
    return torch.cat([ang.sin(), ang.cos()], dim=-1).contiguous()


@pytest.mark.parametrize("half", [False, True])
def test_rope_kernel_forward_and_transpose(half):
    from b200timm import ops
    B, N, H, npt = 3, 37, 4, 2
    g = torch.Generator().manual_seed(1)
    qkv = torch.randn(B * N, 3 * H * 64, generator=g).bfloat16()
    emb = _emb(N - npt) if not half else torch.cat([torch.rand(N - npt, 64, generator=g), torch.rand(N - npt, 64, generator=g)], -1)
    ref = qkv.float().view(B, N, 3, H, 64).clone()
    for which in (0, 1):
        t = ref[:, :, which].permute(0, 2, 1, 3)                # [B, H, N, 64]
        t[:, :, npt:] = O.rope_rotate(t[:, :, npt:].clone(), emb, half)
    out = ops.rope_qk(qkv.cuda().clone(), emb.cuda(), B, N, H, npt, half=half)
    err = (out.float().cpu().view(B, N, 3, H, 64) - ref).abs().max().item()
    assert err <= 2e-2, err                                     # one bf16 rounding of values up to ~4
    assert torch.equal(out.view(B, N, 3, H, 64)[:, :, 2].cpu(), qkv.view(B, N, 3, H, 64)[:, :, 2])      # V untouched
    assert torch.equal(out.view(B, N, 3, H, 64)[:, :npt].cpu(), qkv.view(B, N, 3, H, 64)[:, :npt])      # prefix untouched
    # the inverse flag is the TRANSPOSE of the rotation: <R x, g> == <x, R^T g>
    x = torch.randn(B * N, 3 * H * 64, generator=g).bfloat16()
    gy = torch.randn(B * N, 3 * H * 64, generator=g).bfloat16()
    rx = ops.rope_qk(x.cuda().clone(), emb.cuda(), B, N, H, npt, half=half).float()
    rtg = ops.rope_qk(gy.cuda().clone(), emb.cuda(), B, N, H, npt, half=half, inverse=True).float()
    lhs, rhs = (rx * gy.cuda().float()).sum().item(), (x.cuda().float() * rtg).sum().item()
    assert abs(lhs - rhs) <= 2e-2 * max(abs(lhs), abs(rhs), 1.0) + 5.0, (lhs, rhs)


def test_fused_qkv_rope_epilogue_matches_projection_then_rotation():
    """b200_gemm(EPI_BIAS_ROPE) == nn.Linear (bf16 output) followed by the fp32 rotation and a bf16 rounding -- the
    reference's order of roundings (attention.py:235,252-253)."""
    from b200timm import ops
    B, N, H, npt, D = 5, 197, 12, 1, 768
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B * N, D, generator=g).bfloat16().cuda()
    w = (torch.randn(3 * H * 64, D, generator=g) * 0.05).bfloat16().cuda()
    b = torch.randn(3 * H * 64, generator=g).bfloat16().cuda()
    emb = _emb(N - npt, seed=3).cuda()
    fused = ops.linear_fwd(x, w, b, epilogue=ops.EPI_BIAS_ROPE, rope=(emb, N, npt, 2 * H * 64))
    two_step = ops.rope_qk(ops.linear_fwd(x, w, b), emb, B, N, H, npt)
    assert torch.equal(fused, two_step), (fused.float() - two_step.float()).abs().max()


@pytest.mark.parametrize("layout", ["interleaved", "half"])
def test_attention_rope_module_forward_backward(layout):
    """The AttentionRope mirror in bf16 against an fp32 evaluation of the oracle on the same bf16-rounded weights and
    inputs (golden fixture shapes: 2 prefix tokens, 6 x 5 grid): output and all gradients."""
    from b200timm.layers import AttentionRope
    fx = torch.load(os.path.join(GOLDEN, "rope.pt"))[layout]
    half = layout == "half"
    sd = {k: v.bfloat16().float() for k, v in fx["state_dict"].items()}
    m = AttentionRope(128, num_heads=fx["num_heads"], qkv_bias=True, num_prefix_tokens=fx["num_prefix"], rotate_half=half)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().to(torch.bfloat16).train()
    x = fx["x"].bfloat16()
    xg = x.cuda().requires_grad_(True)
    y = m(xg, rope=fx["emb"])
    (y.float() * fx["wgt"].cuda()).sum().backward()
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.float().requires_grad_(True)
    yr = O.attention_rope(xr, ref_sd["qkv.weight"], ref_sd["qkv.bias"], ref_sd["proj.weight"], ref_sd["proj.bias"],
                          fx["num_heads"], fx["emb"], num_prefix=fx["num_prefix"], half=half)
    (yr * fx["wgt"]).sum().backward()
    assert (y.float().cpu() - yr).abs().max().item() <= 2e-2 + 2e-2 * yr.abs().max().item()

    def close(got, ref, name):
        rel = (got.float().cpu() - ref).norm().item() / max(ref.norm().item(), 1e-6)
        cos = torch.nn.functional.cosine_similarity(got.float().cpu().flatten(), ref.flatten(), dim=0).item()
        assert rel <= 5e-2 and cos >= 0.998, f"{name}: rel L2 err {rel:.4f}, cos {cos:.5f}"

    close(xg.grad, xr.grad, "dx")
    for k, p in m.named_parameters():
        close(p.grad, ref_sd[k].grad, k)
    # inference path (no autograd) gives the same output
    with torch.no_grad():
        y2 = m.eval()(x.cuda(), rope=fx["emb"])
    assert torch.equal(y2, y.detach())


def test_eva_attention_through_the_rope_path():
    """EvaAttention's parameter layout (q_bias, v_bias, zero k_bias buffer, bias-free qkv Linear; eva.py:157-165) runs
    through the same fused path; gradients reach q_bias / v_bias through the per-call concatenation."""
    import torch.nn as nn
    from b200timm import functional as BF
    fx = torch.load(os.path.join(GOLDEN, "rope.pt"))["eva"]

    class EvaAttentionLike(nn.Module):       # attribute-for-attribute what timm.models.eva.EvaAttention holds
        def __init__(self, dim, heads, npt):
            super().__init__()
            self.num_heads, self.head_dim, self.scale, self.num_prefix_tokens = heads, dim // heads, (dim // heads) ** -0.5, npt
            self.rotate_half = False
            self.qkv = nn.Linear(dim, 3 * dim, bias=False)
            self.q_bias = nn.Parameter(torch.zeros(dim))
            self.register_buffer("k_bias", torch.zeros(dim), persistent=False)
            self.v_bias = nn.Parameter(torch.zeros(dim))
            self.q_norm, self.k_norm, self.norm = nn.Identity(), nn.Identity(), nn.Identity()
            self.attn_drop, self.proj_drop = nn.Dropout(0.), nn.Dropout(0.)
            self.proj = nn.Linear(dim, dim)

    sd = {k: v.bfloat16().float() for k, v in fx["state_dict"].items()}
    m = EvaAttentionLike(128, fx["num_heads"], fx["num_prefix"])
    m.load_state_dict(sd, strict=True)
    m = m.cuda().to(torch.bfloat16).train()
    x = fx["x"].bfloat16()
    xg = x.cuda().requires_grad_(True)
    y = BF.attention_rope_forward(m, xg, fx["emb"])
    (y.float() * fx["wgt"].cuda()).sum().backward()
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.float().requires_grad_(True)
    bias = torch.cat((ref_sd["q_bias"], torch.zeros_like(ref_sd["q_bias"]), ref_sd["v_bias"]))
    yr = O.attention_rope(xr, ref_sd["qkv.weight"], bias, ref_sd["proj.weight"], ref_sd["proj.bias"], fx["num_heads"],
                          fx["emb"], num_prefix=fx["num_prefix"])
    (yr * fx["wgt"]).sum().backward()
    assert (y.float().cpu() - yr).abs().max().item() <= 2e-2 + 2e-2 * yr.abs().max().item()
    for k, p in m.named_parameters():
        ref = ref_sd[k].grad
        rel = (p.grad.float().cpu() - ref).norm().item() / max(ref.norm().item(), 1e-6)
        assert rel <= 5e-2, f"{k}: rel L2 err {rel:.4f}"
