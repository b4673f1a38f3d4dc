"""Host-side logic that needs no GPU: module / state_dict compatibility with the reference, the model-name
resolver, mask -> key-length conversion, and the no-CPU-fallback rule."""
import glob
import os

import pytest
import torch

from b200timm import B200Error
from b200timm.layers import Attention, BF_mask_to_kv_len, DropPath, LayerNorm, LayerScale, Mlp, PatchEmbed
from b200timm.vision_transformer import MODEL_ARGS, VisionTransformer, calculate_drop_path_rates, create_model

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "vit_*.pt"))))
def test_state_dict_keys_and_shapes_match_timm(path):
    fx = torch.load(path)
    model = VisionTransformer(**fx["kwargs"])
    ours = model.state_dict()
    ref = fx["state_dict"]
    assert list(ours.keys()) == list(ref.keys())  # same names, same registration order as timm
    for k in ref:
        assert ours[k].shape == ref[k].shape, k
    model.load_state_dict(ref, strict=True)
    if "grads" in fx:  # every parameter timm trains is a parameter here too
        assert sorted(fx["grads"]) == sorted(n for n, _ in model.named_parameters())


def test_vit_base_parameter_count_matches_reference_csv():
    # results/benchmark-*.csv: vit_base_patch16_224 param_count 86.57 M; vit_tiny 5.72 M
    n = sum(p.numel() for p in create_model("vit_base_patch16_224").parameters())
    assert abs(n / 1e6 - 86.57) < 0.01
    n = sum(p.numel() for p in create_model("vit_tiny_patch16_224").parameters())
    assert abs(n / 1e6 - 5.72) < 0.01
    with pytest.raises(B200Error):
        create_model("resnet50")
    assert "vit_large_patch14_clip_336" in MODEL_ARGS


def test_block_children_names_match_reference():
    blk = create_model("vit_tiny_patch16_224", init_values=1e-5, drop_path_rate=0.1).blocks[-1]
    assert [n for n, _ in blk.named_children()] == ["norm1", "attn", "ls1", "drop_path1", "norm2", "mlp", "ls2", "drop_path2"]
    assert [n for n, _ in blk.attn.named_children()] == ["qkv", "q_norm", "k_norm", "attn_drop", "norm", "proj", "proj_drop"]
    assert [n for n, _ in blk.mlp.named_children()] == ["fc1", "act", "drop1", "norm", "fc2", "drop2"]
    assert isinstance(blk.ls1, LayerScale) and isinstance(blk.drop_path1, DropPath)
    assert abs(blk.drop_path1.drop_prob - 0.1) < 1e-6
    assert calculate_drop_path_rates(0.1, 3) == pytest.approx([0.0, 0.05, 0.1])


def test_default_init_statistics():
    m = create_model("vit_tiny_patch16_224")
    w = m.blocks[0].attn.qkv.weight
    assert abs(w.std().item() - 0.02) < 1e-3   # trunc_normal_(std=.02, a=-2, b=2)
    assert m.blocks[0].attn.qkv.bias.abs().max().item() == 0
    assert isinstance(m.norm, LayerNorm) and m.norm.eps == 1e-6
    assert create_model("vit_large_patch14_clip_336", depth=1).norm.eps == 1e-5


def test_mask_to_kv_len():
    B, N = 3, 10
    lens = torch.tensor([10, 4, 7])
    valid = torch.arange(N)[None, :] < lens[:, None]
    keep = valid[:, None, :, None] & valid[:, None, None, :]
    add = torch.zeros(B, 1, N, N).masked_fill(~keep, torch.finfo(torch.float32).min)
    assert BF_mask_to_kv_len(keep, N).tolist() == lens.tolist()
    assert BF_mask_to_kv_len(add, N).tolist() == lens.tolist()
    holes = keep.clone()
    holes[1, 0, :, 1] = False  # a non-prefix mask must be rejected, not silently mis-handled
    with pytest.raises(B200Error):
        BF_mask_to_kv_len(holes, N)


def test_no_cpu_fallback():
    x = torch.zeros(2, 5, 128, dtype=torch.bfloat16)
    for mod in (LayerNorm(128), Mlp(128, 512), Attention(128, num_heads=2)):
        with pytest.raises(B200Error):
            mod.to(torch.bfloat16)(x)
    with pytest.raises(B200Error):
        PatchEmbed(32, 16, 3, 64).to(torch.bfloat16)(torch.zeros(1, 3, 32, 32, dtype=torch.bfloat16))
    with pytest.raises(B200Error):
        Attention(96, num_heads=2).to(torch.bfloat16)(torch.zeros(1, 4, 96, dtype=torch.bfloat16))  # head_dim 48


def test_convnext_state_dict_matches_timm():
    from b200timm.convnext import ConvNeXt, create_model as create_convnext
    fx = torch.load(os.path.join(GOLDEN, "convnext_g1.pt"))
    model = ConvNeXt(**fx["kwargs"])
    assert list(model.state_dict().keys()) == list(fx["state_dict"].keys())
    model.load_state_dict(fx["state_dict"], strict=True)
    fx2 = torch.load(os.path.join(GOLDEN, "convnext_g2_v2_grn.pt"))         # ConvNeXt-V2: mlp.grn.{weight,bias}, no gamma
    m2 = ConvNeXt(**fx2["kwargs"])
    assert list(m2.state_dict().keys()) == list(fx2["state_dict"].keys())
    m2.load_state_dict(fx2["state_dict"], strict=True)
    n = sum(p.numel() for p in create_convnext("convnext_base").parameters())
    assert abs(n / 1e6 - 88.59) < 0.01     # results/benchmark-*.csv param_count for convnext_base
    with pytest.raises(B200Error):
        model.to(torch.bfloat16)(torch.zeros(1, 3, 64, 64, dtype=torch.bfloat16))   # CPU: no fallback


def test_naflex_state_dict_matches_timm():
    from b200timm.naflexvit import NaFlexVit, create_model as create_naflex
    fx = torch.load(os.path.join(GOLDEN, "naflex_g1.pt"))
    model = NaFlexVit(**fx["kwargs"])
    assert sorted(model.state_dict().keys()) == sorted(fx["state_dict"].keys())
    model.load_state_dict(fx["state_dict"], strict=True)
    big = create_naflex("vit_large_patch14_clip_336", depth=1)
    assert big.embeds.pos_embed.shape == (1, 24, 24, 1024) and big.embeds.proj.bias is None and big.norm.eps == 1e-5


def test_optimizer_mirrors_torch_optim_interface_and_has_no_cpu_path():
    """b200timm.optim.{SGD, AdamW, Adam}: torch.optim constructor signatures / defaults / param-group keys (what timm's
    create_optimizer_v2 passes, timm/optim/_optim_factory.py:355-404), argument validation, and no CPU fallback."""
    import inspect
    from b200timm import optim as BO
    for mine, ref in ((BO.SGD, torch.optim.SGD), (BO.AdamW, torch.optim.AdamW), (BO.Adam, torch.optim.Adam)):
        a, b = inspect.signature(mine.__init__).parameters, inspect.signature(ref.__init__).parameters
        for name in ("params", "lr", "weight_decay") + (("momentum", "dampening", "nesterov") if mine is BO.SGD
                                                         else ("betas", "eps", "amsgrad")):
            assert name in a and name in b, name
            if name not in ("params",) and not (mine is BO.Adam and name == "weight_decay"):
                assert a[name].default == b[name].default, (mine.__name__, name)
    p = torch.nn.Parameter(torch.randn(4))
    p.grad = torch.randn(4)
    o = BO.AdamW([dict(params=[p], lr=3e-4)], weight_decay=0.05)
    assert o.param_groups[0]["lr"] == 3e-4 and o.param_groups[0]["weight_decay"] == 0.05
    assert o.param_groups[0]["betas"] == (0.9, 0.999)
    with pytest.raises(B200Error):
        o.step()                                   # CPU parameter: no fallback
    with pytest.raises(B200Error):
        BO.SGD([p], lr=0.1).step()
    with pytest.raises(B200Error):
        BO.clip_grad_norm_([p], 1.0)
    with pytest.raises(B200Error):
        BO.AdamW([p], amsgrad=True)
    with pytest.raises(ValueError):
        BO.SGD([p], lr=0.1, nesterov=True)         # Nesterov needs a momentum, as in torch
    with pytest.raises(ValueError):
        BO.AdamW([p], betas=(1.0, 0.999))


def test_prefetch_loader_mirror_host_logic():
    """b200timm.data.PrefetchLoader: constructor arguments of timm's PrefetchLoader (timm/data/loader.py:82-96), channel
    adaptation of mean / std (:66-76), pass-through properties, and the refusals (random erasing, CPU device)."""
    from b200timm import data as BD, ops

    class _L(list):
        sampler, dataset, collate_fn = "sampler", "dataset", None
    inner = _L([(torch.zeros(2, 1, 8, 8, dtype=torch.uint8), torch.zeros(2, dtype=torch.long))])
    try:
        ld = BD.PrefetchLoader(inner, mean=(0.5, 0.4, 0.3), std=0.25, channels=1, device=torch.device("cpu"))
        assert ld.mean == (pytest.approx(0.4),) and ld.std == (0.25,)
        assert len(ld) == 1 and ld.sampler == "sampler" and ld.dataset == "dataset" and ld.mixup_enabled is False
        norm = ops.input_norm()
        assert [round(v, 3) for v in norm[0]] == [102.0] and list(norm[1]) == [63.75]
        with pytest.raises(B200Error):
            next(iter(ld))                       # CPU device: no fallback
        with pytest.raises(B200Error):
            BD.PrefetchLoader(inner, re_prob=0.25)
        with pytest.raises(B200Error):
            ops.set_input_norm((0.5,), (0.0,))
    finally:
        ops.set_input_norm(None)


def test_loss_and_ema_mirrors_host_logic():
    """b200timm.loss / b200timm.ema: class names and constructor arguments of the reference (timm/loss/cross_entropy.py,
    timm/utils/model_ema.py:156-203), the decay schedule, and the refusal of CPU tensors (no fallback)."""
    from b200timm import ema as BE, loss as BL
    ls = BL.LabelSmoothingCrossEntropy(smoothing=0.2)
    assert ls.smoothing == 0.2 and ls.confidence == 0.8
    with pytest.raises(B200Error):
        BL.CrossEntropyLoss(reduction="sum")
    with pytest.raises(B200Error):
        BL.CrossEntropyLoss(weight=torch.ones(3))
    for mod, tgt in ((BL.CrossEntropyLoss(), torch.tensor([0, 1])), (ls, torch.tensor([0, 1])),
                     (BL.SoftTargetCrossEntropy(), torch.full((2, 5), 0.2))):
        with pytest.raises(B200Error):
            mod(torch.randn(2, 5), tgt)
    with pytest.raises(B200Error):
        BL.SoftTargetCrossEntropy()(torch.randn(2, 5), torch.tensor([0, 1]))
    m = torch.nn.Linear(4, 3)
    e = BE.ModelEma(m, decay=0.998, use_warmup=True, update_after_step=1)
    assert e.module is not m and not e.module.training
    assert [e.get_decay(s) for s in (None, 1, 2)] == [0.998, 0.0, 0.0]
    assert e.get_decay(3) == pytest.approx(1 - 2 ** (-2 / 3)) and e.get_decay(10 ** 9) == 0.998
    e.set(m)
    with pytest.raises(B200Error):
        e.update(m, step=5)                       # CPU model: no fallback
    e32 = BE.ModelEma(m.to(torch.bfloat16), ema_dtype=torch.float32)
    assert e32.module.weight.dtype == torch.float32


@pytest.mark.parametrize("D,H,hd,P", [(96, 3, 32, 64), (160, 2, 80, 128), (144, 2, 72, 128), (192, 2, 96, 128)])
def test_head_padding_in_the_weights_is_exact(D, H, hd, P):
    """head_dim other than 64 / 128 runs on the 64- / 128-wide kernels through zero padding that lives in the weights
    (functional._pad_rows / _pad_cols): attention computed on the padded projections with the module's own scale equals
    the narrow-head attention exactly (the padding columns of q / k add 0 to every score, the padded v columns meet zero
    columns of the output projection).  Checked here with plain torch math (the algebra, not the kernels)."""
    from b200timm import functional as BF
    torch.manual_seed(0)
    B, N = 2, 7
    attn = Attention(D, num_heads=H, qkv_bias=True)
    x = torch.randn(B, N, D)

    def sdpa(qkv, heads, width, scale):
        q, k, v = qkv.view(B, N, 3, heads, width).permute(2, 0, 3, 1, 4)
        p = ((q @ k.transpose(-1, -2)) * scale).softmax(-1)
        return (p @ v).transpose(1, 2).reshape(B, N, heads * width)

    ref = torch.nn.functional.linear(sdpa(torch.nn.functional.linear(x, attn.qkv.weight, attn.qkv.bias), H, hd, attn.scale),
                                     attn.proj.weight, attn.proj.bias)
    wq = BF._pad_rows(attn.qkv.weight, 3, H, hd)
    bq = BF._pad_rows(attn.qkv.bias, 3, H, hd)
    wp = BF._pad_cols(attn.proj.weight, H, hd)
    assert wq.shape == (3 * H * P, D) and bq.shape == (3 * H * P,) and wp.shape == (D, H * P)
    got = torch.nn.functional.linear(sdpa(torch.nn.functional.linear(x, wq, bq), H, P, attn.scale), wp, attn.proj.bias)
    torch.testing.assert_close(got, ref, atol=1e-6, rtol=1e-6)
    # gradients reach the real parameters through the padding (training builds the padded operands differentiably)
    got.sum().backward()
    assert attn.qkv.weight.grad is not None and attn.proj.weight.grad is not None and attn.qkv.bias.grad is not None
    assert BF.head_pad(32) == 64 and BF.head_pad(64) == 64 and BF.head_pad(80) == 128 and BF.head_pad(128) == 128
    for bad in (136, 36, 0):
        with pytest.raises(B200Error):
            BF.head_pad(bad)


def test_attn_mask_operand_classifies_masks():
    """layers.attn_mask_operand: (attn_mask, is_causal) of the reference signatures (attention.py:21-40) -> None / int32
    valid-key counts for prefix key-padding masks (the NaFlex layout) / an fp32 additive bias for everything else."""
    import b200timm.layers as L
    from b200timm._lib import B200Error
    N, B = 6, 2
    assert L.attn_mask_operand(None, False, N, B) is None
    c = L.attn_mask_operand(None, True, N, B, torch.device("cpu"))
    assert c.shape == (1, 1, N, 8) and c.dtype == torch.float32
    assert torch.equal(c[0, 0, :, :N], torch.full((N, N), float("-inf")).triu_(1)) and bool((c[..., N:] == 0).all())
    assert L.attn_mask_operand(torch.ones(B, 1, N, N, dtype=torch.bool), True, N, B, torch.device("cpu")) is c   # causal wins
    # NaFlex: q_valid & k_valid, bool and additive (finfo.min) forms -> key counts
    valid = torch.tensor([[1, 1, 1, 1, 0, 0], [1, 1, 1, 1, 1, 1]], dtype=torch.bool)
    m = (valid[:, None, :, None] & valid[:, None, None, :])
    kv = L.attn_mask_operand(m, False, N, B)
    assert kv.dtype == torch.int32 and kv.tolist() == [4, 6]
    add = torch.zeros(B, 1, N, N).masked_fill(~m, torch.finfo(torch.float32).min)
    assert L.attn_mask_operand(add, False, N, B).tolist() == [4, 6]
    assert L.attn_mask_operand(m, False, N, B) is kv                      # cached on the mask object
    # a lower-triangular keep-mask keeps every key for SOME query but is not key padding -> additive bias
    tri = torch.ones(N, N, dtype=torch.bool).tril()
    t = L.attn_mask_operand(tri, False, N, B)
    assert t.dtype == torch.float32 and t.shape == (1, 1, N, 8)
    assert torch.equal(t[0, 0, :, :N], torch.zeros(N, N).masked_fill(~tri, float("-inf")))
    with pytest.raises(B200Error):
        L.BF_mask_to_kv_len(tri, N, B)                                    # strict callers still reject it
    # float bias per head; key-only [B, 1, 1, N] masks broadcast over the queries
    fb = torch.randn(1, 3, N, N)
    f = L.attn_mask_operand(fb, False, N, B)
    assert f.shape == (1, 3, N, 8) and torch.equal(f[..., :N], fb)
    hole = torch.tensor([[1, 0, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1]], dtype=torch.bool)[:, None, None, :]
    h = L.attn_mask_operand(hole, False, N, B)
    assert h.shape == (B, 1, N, 8) and bool(torch.isinf(h[0, 0, :, 1]).all()) and bool((h[1] == 0).all())
    with pytest.raises(B200Error):
        L.attn_bias_operand(torch.zeros(2, 1, N, N + 1), N)


def test_folded_layernorm_operands_are_the_same_affine_map():
    """functional._ln_fold: Linear(LayerNorm(x)) == rstd * (x W'^T - mean * c1) + c2 with W' = W * gamma, c1 = rowsum(W'),
    c2 = W beta + b (timm/layers/norm.py:85-90 followed by attention.py:115 / mlp.py:48) -- the algebra the GEMM epilogue
    evaluates (b200_gemm_desc.ln_*), checked in fp64 torch math on the operands the host builds; the cache follows the
    parameters' version counters."""
    from b200timm import functional as BF
    torch.manual_seed(0)
    D, N = 192, 384
    norm, lin = torch.nn.LayerNorm(D, eps=1e-6), torch.nn.Linear(D, N)
    with torch.no_grad():
        norm.weight.normal_(1.0, 0.3)
        norm.bias.normal_(0.0, 0.3)
    norm, lin = norm.to(torch.bfloat16), lin.to(torch.bfloat16)
    f = BF._ln_fold(norm, lin, lin.weight, lin.bias, "t")
    assert f is not None and f is not BF._ln_fold(norm, lin, lin.weight, lin.bias, "t")       # tuples; the cached tensors are shared
    wg, c1, c2, eps = f
    assert wg.dtype == torch.bfloat16 and c1.dtype == c2.dtype == torch.float32 and eps == 1e-6
    assert BF._ln_fold(norm, lin, lin.weight, lin.bias, "t")[0] is wg
    x = (torch.randn(17, D) * 2 + 0.7).bfloat16().double()
    mean = x.mean(-1, keepdim=True)
    rstd = (x.var(-1, unbiased=False, keepdim=True) + eps).rsqrt()
    got = rstd * (x @ wg.double().t() - mean * c1.double()) + c2.double()
    xh = (x - mean) * rstd
    # reference with the SAME rounded operand W' (what the tensor cores multiply): exact up to fp64 rounding
    ref_same = xh @ wg.double().t() + c2.double()
    torch.testing.assert_close(got, ref_same, atol=1e-5, rtol=1e-6)      # c1 is held in fp32
    # and against the module pair itself: differs only by the bf16 rounding of W * gamma
    ref = torch.nn.functional.linear(torch.nn.functional.layer_norm(x, (D,), norm.weight.double(), norm.bias.double(), eps),
                                     lin.weight.double(), lin.bias.double())
    assert (got - ref).abs().max().item() < 2e-2 * ref.abs().max().item()
    with torch.no_grad():
        norm.weight.mul_(2.0)                     # version bump -> rebuilt
    assert BF._ln_fold(norm, lin, lin.weight, lin.bias, "t")[0] is not wg
    assert BF._ln_fold(torch.nn.Identity(), lin, lin.weight, lin.bias, "t") is None
    small = torch.nn.Linear(D, 64).to(torch.bfloat16)
    assert BF._ln_fold(norm, small, small.weight, small.bias, "t") is None       # N <= 128: no folded instantiation
    carry = BF.LnCarry()
    assert carry.stats is None


def test_weight_gradient_row_chunk_plan(monkeypatch):
    """ops._wgrad_parts (host logic): which weight gradients run as several launches over row chunks on a 148-SM device."""
    from b200timm import ops

    class Props:
        multi_processor_count = 148
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda d: Props())
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    assert ops._wgrad_parts(3072, 1024, 36928) == [(0, 1536, 3), (1536, 1536, 3)]      # ViT-L qkv: 48 tiles -> 2 x 24 x split 3
    for n, k in ((2304, 768), (768, 768), (3072, 768), (768, 3072), (1024, 1024), (4096, 1024), (1152, 384)):
        assert ops._wgrad_parts(n, k, 50432) is None, (n, k)                             # fill the pairs, or would need a deep split
    assert ops._wgrad_parts(3072, 1024, 128) is None                                     # short K: nothing to split
    monkeypatch.setattr(ops, "_WGRAD_NSPLIT", False)
    assert ops._wgrad_parts(3072, 1024, 36928) is None
