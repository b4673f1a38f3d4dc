"""Attention heads wider than 64 (timm/layers/attention.py:85-93: head_dim = attn_head_dim or dim // num_heads -- 72 in SigLIP
SO400M, 80 in ViT-H/14, 96, 128): the 128-wide instantiations of the long-sequence kernels (attention_fwd_long.cu /
attention_bwd_long.cu, NH = 2) against fp32 autograd of the restated op, and whole models whose 72 / 80 / 96-wide heads reach
those kernels zero-padded through the weights (functional.head_pad) against the CPU oracle.

Tolerances as in test_gpu_kernels.py / test_gpu_train.py (bf16 outputs, fp32 accumulation)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gpu_kernels import _attn_ref, _close, _ops      # noqa: E402


@pytest.mark.parametrize("B,N,H", [(2, 197, 3), (3, 64, 2), (1, 17, 1), (2, 257, 4), (1, 577, 2), (2, 129, 1), (1, 1025, 1)])
def test_attention_128_wide_heads_fwd_bwd(B, N, H):
    ops = _ops()
    hd = 128
    g = torch.Generator(device=DEV).manual_seed(31 * N + H)
    qkv = torch.randn(B, N, 3 * H * hd, device=DEV, generator=g).bfloat16()
    dout = torch.randn(B, N, H * hd, device=DEV, generator=g).bfloat16()
    scale = hd ** -0.5
    out, lse = ops.attention_fwd(qkv, B, N, H, hd, scale, need_lse=True)
    ref_in = qkv.float().requires_grad_(True)
    ref, lse_ref = _attn_ref(ref_in, B, N, H, hd, scale)
    _close(out, ref, 1e-2, 2e-2, f"attention hd=128 B={B} N={N} H={H}")
    _close(lse, lse_ref, 1e-2, 1e-3, "lse")
    if N <= 1120:
        dqkv = ops.attention_bwd(qkv, out, dout, lse, B, N, H, hd, scale)
        ref.backward(dout.float())
        _close(dqkv, ref_in.grad, 2e-2, 2e-2, f"attention bwd hd=128 B={B} N={N} H={H}")


@pytest.mark.parametrize("N", [197, 300])
def test_attention_128_wide_heads_key_padding_and_bias(N):
    import b200timm.layers as L
    ops = _ops()
    B, H, hd = 4, 2, 128
    g = torch.Generator(device=DEV).manual_seed(N)
    qkv = torch.randn(B, N, 3 * H * hd, device=DEV, generator=g).bfloat16()
    dout = torch.randn(B, N, H * hd, device=DEV, generator=g).bfloat16()
    scale = hd ** -0.5
    kv_len = torch.tensor([N, 1, N // 2, 130], device=DEV, dtype=torch.int32)
    out, lse = ops.attention_fwd(qkv, B, N, H, hd, scale, kv_len=kv_len, need_lse=True)
    dqkv = ops.attention_bwd(qkv, out, dout, lse, B, N, H, hd, scale, kv_len=kv_len)
    ref_in = qkv.float().requires_grad_(True)
    ref, lse_ref = _attn_ref(ref_in, B, N, H, hd, scale, kv_len)
    ref.backward(dout.float())
    _close(out, ref, 1e-2, 2e-2, "masked attention hd=128")
    _close(lse, lse_ref, 1e-2, 1e-3, "masked lse hd=128")
    _close(dqkv, ref_in.grad, 2e-2, 2e-2, "masked attention bwd hd=128")
    # causal triangle through the additive-bias variants
    operand = L.attn_mask_operand(None, True, N, B, DEV)
    out, lse = ops.attention_fwd(qkv, B, N, H, hd, scale, kv_len=operand, need_lse=True)
    dqkv = ops.attention_bwd(qkv, out, dout, lse, B, N, H, hd, scale, kv_len=operand)
    ref_in = qkv.float().requires_grad_(True)
    q, k, v = ref_in.view(B, N, 3, H, hd).permute(2, 0, 3, 1, 4).unbind(0)
    sc = (q @ k.transpose(-1, -2)) * scale + torch.full((N, N), float("-inf"), device=DEV).triu_(1)
    ref = (sc.softmax(-1) @ v).transpose(1, 2).reshape(B, N, H * hd)
    ref.backward(dout.float())
    _close(out, ref, 1e-2, 2e-2, "causal attention hd=128")
    _close(dqkv, ref_in.grad, 2e-2, 2e-2, "causal attention bwd hd=128")


def test_attention_rejects_unpadded_widths():
    ops = _ops()
    qkv = torch.zeros(1, 16, 3 * 80, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(ops.B200Error):
        ops.attention_fwd(qkv, 1, 16, 1, 80, 0.1)


@pytest.mark.parametrize("embed_dim,num_heads,img", [(160, 2, 96), (144, 2, 64), (192, 2, 224), (256, 2, 96)])   # 80, 72, 96, 128
def test_vit_with_wide_heads_matches_oracle(embed_dim, num_heads, img):
    """Eval logits and every parameter gradient of a ViT whose head_dim is 80 / 72 / 96 / 128 against the CPU oracle."""
    from oracle import vit_oracle as O
    from b200timm.vision_transformer import VisionTransformer
    from test_gpu_train import _check_grads
    cfg = dict(img_size=img, patch_size=16, embed_dim=embed_dim, depth=2, num_heads=num_heads)
    sd = {k: v.bfloat16().float() for k, v in O.random_state_dict(cfg, seed=11, num_classes=24).items()}
    model = VisionTransformer(num_classes=24, **cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().to(torch.bfloat16)
    g = torch.Generator().manual_seed(3)
    batch = 3
    x = torch.randn(batch, 3, img, img, generator=g).bfloat16()
    target = torch.randint(0, 24, (batch,), generator=g)
    ocfg = dict(cfg, num_prefix_tokens=1)
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = O.vit_forward(x.float(), ref_sd, ocfg)
    model.eval()
    with torch.no_grad():
        got = model(x.cuda()).float().cpu()
    assert (got - ref.detach()).abs().max().item() <= 2e-2 + 2e-2 * ref.abs().max().item()
    model.train()
    loss = torch.nn.functional.cross_entropy(model(x.cuda()).float(), target.cuda())
    loss.backward()
    ref_loss = torch.nn.functional.cross_entropy(ref, target)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 2e-2
    _check_grads(model, {k: v.grad for k, v in ref_sd.items()})
