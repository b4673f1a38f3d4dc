"""Pins the oracle restatement (oracle/vit_oracle.py) against golden vectors produced by the unmodified
reference (timm) -- see oracle/make_golden.py.  CPU only; fp32; tolerance 1e-5 absolute (the reference's own
fused-vs-unfused attention check uses atol=1e-5, tests/test_layers.py:141-160)."""
import glob
import os

import pytest
import torch

from oracle import vit_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VIT_FILES = sorted(glob.glob(os.path.join(GOLDEN, "vit_*.pt")))


def _cfg(kwargs):
    cfg = dict(kwargs)
    cfg["num_prefix_tokens"] = (1 if kwargs.get("class_token", True) else 0) + kwargs.get("reg_tokens", 0)
    return cfg


def test_golden_files_present():
    assert len(VIT_FILES) >= 3 and os.path.exists(os.path.join(GOLDEN, "layers.pt"))


@pytest.mark.parametrize("path", VIT_FILES, ids=[os.path.basename(p)[:-3] for p in VIT_FILES])
def test_vit_forward_matches_reference(path):
    fx = torch.load(path)
    cfg = _cfg(fx["kwargs"])
    feats, blocks = O.vit_forward_features(fx["img"], fx["state_dict"], cfg, return_blocks=True)
    for i, (got, ref) in enumerate(zip(blocks, fx["block_outputs"])):
        torch.testing.assert_close(got, ref, atol=1e-5, rtol=1e-5, msg=lambda m: f"block {i}: {m}")
    torch.testing.assert_close(feats, fx["features"], atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(O.vit_forward_head(feats, fx["state_dict"], cfg), fx["logits"], atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("path", VIT_FILES, ids=[os.path.basename(p)[:-3] for p in VIT_FILES])
def test_eager_timing_port_matches_reference(path):
    """The CPU-baseline port (vit_forward_eager, timed by bench.py) computes the same function."""
    fx = torch.load(path)
    cfg = _cfg(fx["kwargs"])
    with torch.no_grad():
        torch.testing.assert_close(O.vit_forward_eager(fx["img"], fx["state_dict"], cfg), fx["logits"], atol=1e-5, rtol=1e-5)
        kv = torch.full((fx["img"].shape[0],), 3, dtype=torch.int64)
        got = O.vit_forward_eager(fx["img"], fx["state_dict"], cfg, kv_len=kv)
        ref = O.vit_forward(fx["img"], fx["state_dict"], cfg, kv_len=kv)
    if cfg.get("global_pool", "token") == "token":   # masked queries (avg pool over padding) are unspecified
        torch.testing.assert_close(got, ref, atol=1e-5, rtol=1e-5)


GRAD_FILES = [p for p in VIT_FILES if "grads" in torch.load(p)]


@pytest.mark.parametrize("path", GRAD_FILES, ids=[os.path.basename(p)[:-3] for p in GRAD_FILES])
def test_vit_backward_matches_reference(path):
    """Gradients of every parameter after outputs.mean().backward() (tests/test_models.py:209-242 pattern), for every ViT
    fixture that carries the reference's gradients (default, RMSNorm, SwiGLU, qk-norm, dynamic size, map pool, SigLIP-like).
    Gradients that are analytically zero (a bias on k) hold cancellation noise on both sides: absolute tolerance scaled
    to the largest gradient."""
    fx = torch.load(path)
    cfg = _cfg(fx["kwargs"])
    sd = {k: v.clone().requires_grad_(True) for k, v in fx["state_dict"].items()}
    O.vit_forward(fx["img"], sd, cfg).mean().backward()
    assert set(fx["grads"]) == set(sd)
    scale = max(g.abs().max().item() for g in fx["grads"].values())
    for k, ref in fx["grads"].items():
        torch.testing.assert_close(sd[k].grad, ref, atol=1e-6 + 2e-6 * scale, rtol=2e-4, msg=lambda m: f"{k}: {m}")


def test_wide_head_vit_matches_reference():
    """head_dim 80 (dim // num_heads of ViT-H, attention.py:85-93): the oracle against the unmodified reference's block
    outputs, logits and parameter gradients -- the pin behind tests/test_gpu_wide_heads.py, which checks the 128-wide
    attention kernels (heads padded 80 -> 128 through the weights) against this oracle."""
    fx = torch.load(os.path.join(GOLDEN, "wide_vit_g10_head80.pt"))
    cfg = _cfg(fx["kwargs"])
    assert cfg["embed_dim"] // cfg["num_heads"] == 80
    feats, blocks = O.vit_forward_features(fx["img"], fx["state_dict"], cfg, return_blocks=True)
    for i, (got, ref) in enumerate(zip(blocks, fx["block_outputs"])):
        torch.testing.assert_close(got, ref, atol=1e-5, rtol=1e-5, msg=lambda m: f"block {i}: {m}")
    torch.testing.assert_close(O.vit_forward_head(feats, fx["state_dict"], cfg), fx["logits"], atol=1e-5, rtol=1e-5)
    sd = {k: v.clone().requires_grad_(True) for k, v in fx["state_dict"].items()}
    O.vit_forward(fx["img"], sd, cfg).mean().backward()
    scale = max(g.abs().max().item() for g in fx["grads"].values())
    for k, ref in fx["grads"].items():
        torch.testing.assert_close(sd[k].grad, ref, atol=1e-6 + 2e-6 * scale, rtol=2e-4, msg=lambda m: f"{k}: {m}")


def test_layer_goldens():
    fx = torch.load(os.path.join(GOLDEN, "layers.pt"))
    a = fx["attention"]
    sd = a["state_dict"]
    args = (sd["qkv.weight"], sd["qkv.bias"], sd["proj.weight"], sd["proj.bias"], a["num_heads"])
    torch.testing.assert_close(O.attention(a["x"], *args), a["y"], atol=1e-5, rtol=1e-5)
    # masked: rows whose query is itself padding are unspecified in the reference (all -inf -> uniform);
    # compare the valid queries only, as NaFlex pooling does (naflexvit.py:1033-1102).
    ym = O.attention(a["x"], *args, kv_len=a["kv_len"])
    for b, n in enumerate(a["kv_len"].tolist()):
        torch.testing.assert_close(ym[b, :n], a["y_masked"][b, :n], atol=1e-5, rtol=1e-5)
    m = fx["mlp"]
    sd = m["state_dict"]
    torch.testing.assert_close(O.mlp(m["x"], sd["fc1.weight"], sd["fc1.bias"], sd["fc2.weight"], sd["fc2.bias"]),
                               m["y"], atol=1e-5, rtol=1e-5)
    ln = fx["layernorm"]
    torch.testing.assert_close(O.layer_norm(ln["x"], ln["state_dict"]["weight"], ln["state_dict"]["bias"], ln["eps"]),
                               ln["y"], atol=1e-5, rtol=1e-5)
    for p in (16, 14, 4):
        pe = fx[f"patch_embed_p{p}"]
        torch.testing.assert_close(O.patch_embed(pe["img"], pe["state_dict"]["proj.weight"],
                                                 pe["state_dict"]["proj.bias"], p), pe["y"], atol=1e-5, rtol=1e-5)
    ls = fx["layerscale"]
    torch.testing.assert_close(ls["x"] * ls["state_dict"]["gamma"], ls["y"])
    dp = fx["droppath"]
    torch.testing.assert_close(O.drop_path(dp["x"], dp["keep_mask"], 1.0 - dp["drop_prob"]), dp["y"])


@pytest.mark.parametrize("name", ["convnext_g1", "convnext_g2_v2_grn", "convnext_g3_tanh_k7553"])
def test_convnext_oracle_matches_reference(name):
    from oracle import convnext_oracle as C
    fx = torch.load(os.path.join(GOLDEN, name + ".pt"))
    cfg = dict(fx["kwargs"])
    with torch.no_grad():
        feats, stages = C.convnext_forward_features(fx["img"], fx["state_dict"], cfg, return_stages=True)
        for i, (got, ref) in enumerate(zip(stages, fx["stage_outputs"])):
            torch.testing.assert_close(got, ref, atol=2e-5, rtol=1e-5, msg=lambda m: f"stage {i}: {m}")
        torch.testing.assert_close(feats, fx["features"], atol=2e-5, rtol=1e-5)
        torch.testing.assert_close(C.convnext_forward_head(feats, fx["state_dict"], cfg), fx["logits"], atol=2e-5, rtol=1e-5)
        torch.testing.assert_close(C.convnext_forward_eager(fx["img"], fx["state_dict"], cfg), fx["logits"], atol=2e-5, rtol=1e-5)


def test_naflex_oracle_matches_reference():
    from oracle import naflex_oracle as NO
    fx = torch.load(os.path.join(GOLDEN, "naflex_g1.pt"))
    with torch.no_grad():
        feats, logits = NO.naflex_forward(fx["patches"], fx["patch_coord"], fx["patch_valid"], fx["state_dict"],
                                          dict(fx["kwargs"], eps=1e-6))
    torch.testing.assert_close(logits, fx["logits"], atol=1e-5, rtol=1e-5)
    npre = 1
    for b in range(feats.shape[0]):   # padded query rows are unspecified in the reference (all-masked softmax)
        n = int(fx["patch_valid"][b].sum()) + npre
        torch.testing.assert_close(feats[b, :n], fx["features"][b, :n], atol=1e-5, rtol=1e-5)


def test_naflex_oracle_autograd_matches_reference_grads():
    """Autograd of the oracle restatement reproduces the reference's parameter gradients (outputs.mean().backward()),
    including the position table's gradient through the per-sample bicubic resize."""
    from oracle import naflex_oracle as NO
    fx = torch.load(os.path.join(GOLDEN, "naflex_g1.pt"))
    sd = {k: v.clone().requires_grad_(True) for k, v in fx["state_dict"].items()}
    _, logits = NO.naflex_forward(fx["patches"], fx["patch_coord"], fx["patch_valid"], sd, dict(fx["kwargs"], eps=1e-6))
    logits.mean().backward()
    assert sorted(sd) == sorted(fx["grads"])
    for k, g in fx["grads"].items():
        assert torch.allclose(sd[k].grad, g, atol=1e-6 + 1e-4 * g.abs().max().item(), rtol=1e-3), (
            k, (sd[k].grad - g).abs().max().item(), g.abs().max().item())


@pytest.mark.parametrize("layout", ["interleaved", "half"])
def test_attention_rope_oracle_matches_reference(layout):
    """timm.layers.AttentionRope (attention.py:224-282) with a RotaryEmbeddingCat table: the oracle restatement and its
    autograd reproduce the reference's output, input gradient and every parameter gradient (fp32)."""
    fx = torch.load(os.path.join(GOLDEN, "rope.pt"))[layout]
    sd = {k: v.clone().requires_grad_(True) for k, v in fx["state_dict"].items()}
    x = fx["x"].clone().requires_grad_(True)
    y = O.attention_rope(x, sd["qkv.weight"], sd["qkv.bias"], sd["proj.weight"], sd["proj.bias"], fx["num_heads"],
                         fx["emb"], num_prefix=fx["num_prefix"], half=(layout == "half"))
    assert torch.allclose(y, fx["y"], atol=2e-5, rtol=1e-4), (y - fx["y"]).abs().max()
    (y * fx["wgt"]).sum().backward()
    assert torch.allclose(x.grad, fx["dx"], atol=2e-4, rtol=1e-3), (x.grad - fx["dx"]).abs().max()
    for k, g in fx["grads"].items():
        assert torch.allclose(sd[k].grad, g, atol=5e-4, rtol=1e-3), (k, (sd[k].grad - g).abs().max())


def test_eva_attention_oracle_matches_reference():
    """timm.models.eva.EvaAttention (eva.py:197-262; q_bias / zero k_bias / v_bias concatenated per call): the same oracle
    restatement with the concatenated bias reproduces output and gradients."""
    fx = torch.load(os.path.join(GOLDEN, "rope.pt"))["eva"]
    sd = {k: v.clone().requires_grad_(True) for k, v in fx["state_dict"].items()}
    x = fx["x"].clone().requires_grad_(True)
    bias = torch.cat((sd["q_bias"], torch.zeros_like(sd["q_bias"]), sd["v_bias"]))
    y = O.attention_rope(x, sd["qkv.weight"], bias, sd["proj.weight"], sd["proj.bias"], fx["num_heads"], fx["emb"],
                         num_prefix=fx["num_prefix"])
    assert torch.allclose(y, fx["y"], atol=2e-5, rtol=1e-4), (y - fx["y"]).abs().max()
    (y * fx["wgt"]).sum().backward()
    assert torch.allclose(x.grad, fx["dx"], atol=2e-4, rtol=1e-3)
    for k, g in fx["grads"].items():
        assert torch.allclose(sd[k].grad, g, atol=5e-4, rtol=1e-3), (k, (sd[k].grad - g).abs().max())


def _oracle_optim_run(case, init, grads, round_to=None):
    """Drive oracle/optim_oracle.py through one golden case; returns the parameter trajectory and the clip norms."""
    from oracle import optim_oracle as OO
    kw = dict(case["kwargs"])
    params = [t.clone() if round_to is None else t.to(round_to).float() for t in init]
    s1 = [torch.zeros_like(t) for t in init]
    s2 = [torch.zeros_like(t) for t in init]
    traj, norms = [], []
    for k, gk in enumerate(grads):
        gk = [g if round_to is None else g.to(round_to).float() for g in gk]
        scale = 1.0
        if case["clip"] is not None:
            total, scale = OO.clip_coef(gk, case["clip"])
            norms.append(total)
        if case["opt"] in ("sgd", "momentum"):
            OO.sgd_step(params, gk, s1, kw["lr"], kw.get("momentum", 0.0), kw.get("dampening", 0.0),
                        kw.get("weight_decay", 0.0), nesterov=(case["opt"] == "sgd"), first=(k == 0), grad_scale=scale,
                        round_to=round_to)
        else:
            b1, b2 = kw.get("betas", (0.9, 0.999))
            OO.adam_step(params, gk, s1, s2, k + 1, kw["lr"], b1, b2, kw.get("eps", 1e-8), kw.get("weight_decay", 0.0),
                         decoupled=(case["opt"] == "adamw"), grad_scale=scale, round_to=round_to)
        traj.append([p.clone() for p in params])
    return traj, norms


def test_optimizer_oracle_matches_reference():
    """oracle/optim_oracle.py against the trajectories of timm's create_optimizer_v2 optimizers + dispatch_clip_grad."""
    fx = torch.load(os.path.join(GOLDEN, "optim.pt"))
    assert set(fx["cases"]) >= {"sgd_nesterov_wd", "momentum_dampening", "sgd_plain", "adamw", "adam_l2", "adamw_clip",
                                "sgd_clip"}
    for name, case in fx["cases"].items():
        traj, norms = _oracle_optim_run(case, fx["init"], fx["grads"])
        for k, (got, ref) in enumerate(zip(traj, case["params"])):
            for a, b in zip(got, ref):
                torch.testing.assert_close(a, b, atol=2e-7, rtol=2e-6, msg=lambda m: f"{name} step {k}: {m}")
        for a, b in zip(norms, case["norms"]):
            assert abs(a - b) <= 1e-5 * b, (name, a, b)


def test_ema_oracle_matches_reference():
    """oracle ema_decay / ema_update against timm.utils.ModelEmaV3 with warm-up (tests/golden/ema.pt)."""
    from oracle import optim_oracle as OO
    fx = torch.load(os.path.join(GOLDEN, "ema.pt"))
    ema = {k: v.clone() for k, v in fx["init"].items()}
    for step, (msd, want, d) in enumerate(zip(fx["model_states"], fx["ema_states"], fx["decays"]), start=1):
        decay = OO.ema_decay(step, **fx["kwargs"])
        assert decay == d
        OO.ema_update(ema, msd, decay)
        for k in want:
            torch.testing.assert_close(ema[k], want[k], atol=1e-7, rtol=1e-6, msg=lambda m: f"step {step} {k}: {m}")


def test_loss_oracle_matches_reference():
    """oracle cross_entropy (+ its autograd) against nn.CrossEntropyLoss / timm LabelSmoothingCrossEntropy /
    SoftTargetCrossEntropy values and logits gradients (tests/golden/loss.pt)."""
    from oracle import optim_oracle as OO
    fx = torch.load(os.path.join(GOLDEN, "loss.pt"))
    for name, case in fx["cases"].items():
        for key, tgt, s in (("ce", case["target"], 0.0), ("ls", case["target"], 0.1), ("soft", case["soft_target"], 0.0)):
            x = case["logits"].clone().requires_grad_(True)
            loss = OO.cross_entropy(x, tgt, s)
            (loss * case[key]["grad_out"]).backward()
            torch.testing.assert_close(loss.detach(), case[key]["loss"], atol=1e-6, rtol=1e-6)
            torch.testing.assert_close(x.grad, case[key]["dlogits"], atol=1e-8, rtol=1e-5, msg=lambda m: f"{name}/{key}: {m}")


@pytest.mark.parametrize("name", ["swiglu", "swiglu_norm", "swiglu_packed", "glu_sigmoid"])
def test_glu_mlp_oracle_matches_reference(name):
    """oracle glu_mlp (+ autograd) against timm.layers.SwiGLU / SwiGLUPacked / GluMlp (tests/golden/glu.pt)."""
    fx = torch.load(os.path.join(GOLDEN, "glu.pt"))[name]
    sd = {k: v.clone().requires_grad_(True) for k, v in fx["state_dict"].items()}
    x = fx["x"].clone().requires_grad_(True)
    y = O.glu_mlp(x, sd, "", act={"SiLU": "silu", "Sigmoid": "sigmoid"}[fx["act"]], gate_last=bool(fx["gate_last"]))
    torch.testing.assert_close(y, fx["y"], atol=2e-6, rtol=1e-5)
    (y * fx["wgt"]).sum().backward()
    torch.testing.assert_close(x.grad, fx["dx"], atol=1e-5, rtol=1e-4)
    for k, g in fx["grads"].items():
        torch.testing.assert_close(sd[k].grad, g, atol=1e-4, rtol=1e-4, msg=lambda m: f"{k}: {m}")


def test_swiglu_vit_oracle_matches_reference():
    fx = torch.load(os.path.join(GOLDEN, "vit_g5_swiglu.pt"))
    cfg = dict(fx["kwargs"], eps=1e-6)
    with torch.no_grad():
        feats = O.vit_forward_features(fx["img"], fx["state_dict"], cfg)
        logits = O.vit_forward_head(feats, fx["state_dict"], cfg)
    torch.testing.assert_close(logits, fx["logits"], atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("name", ["ln_qk", "rms_qk_scale", "ln_scale_only"])
def test_attention_norm_oracle_matches_reference(name):
    """oracle attention with q_norm / k_norm / norm (+ autograd) against timm.layers.Attention(qk_norm, scale_norm)."""
    fx = torch.load(os.path.join(GOLDEN, "qknorm.pt"))[name]
    sd = {k: v.clone().requires_grad_(True) for k, v in fx["state_dict"].items()}
    x = fx["x"].clone().requires_grad_(True)
    eps = 1e-6
    y = O.attention(x, sd["qkv.weight"], sd.get("qkv.bias"), sd["proj.weight"], sd["proj.bias"], fx["num_heads"],
                    norms=(sd, "", eps))
    torch.testing.assert_close(y, fx["y"], atol=2e-5, rtol=1e-4)
    (y * fx["wgt"]).sum().backward()
    torch.testing.assert_close(x.grad, fx["dx"], atol=2e-4, rtol=1e-3)
    for k, g in fx["grads"].items():
        torch.testing.assert_close(sd[k].grad, g, atol=5e-4, rtol=1e-3, msg=lambda m: f"{k}: {m}")


@pytest.mark.parametrize("name", ["convnext_g2_v2_grn", "convnext_g3_tanh_k7553"])
def test_convnext_oracle_autograd_matches_reference_grads(name):
    """ConvNeXt-V2 (GlobalResponseNorm in every block) and the tanh-GELU / mixed-kernel-size configuration: autograd of
    the oracle against the reference's parameter gradients."""
    from oracle import convnext_oracle as C
    fx = torch.load(os.path.join(GOLDEN, name + ".pt"))
    sd = {k: v.clone().requires_grad_(True) for k, v in fx["state_dict"].items()}
    C.convnext_forward(fx["img"], sd, dict(fx["kwargs"])).mean().backward()
    assert sorted(sd) == sorted(fx["grads"])
    for k, g in fx["grads"].items():
        assert torch.allclose(sd[k].grad, g, atol=1e-7 + 2e-4 * g.abs().max().item(), rtol=2e-3), (
            k, (sd[k].grad - g).abs().max().item(), g.abs().max().item())



@pytest.mark.parametrize("name", ["gelu_tanh", "silu"])
def test_mlp_activation_oracle_matches_reference(name):
    """oracle mlp(act=...) (+ autograd) against timm.layers.Mlp(act_layer=get_act_layer(name)) (tests/golden/mlp_acts.pt)."""
    fx = torch.load(os.path.join(GOLDEN, "mlp_acts.pt"))[name]
    sd = {k: v.clone().requires_grad_(True) for k, v in fx["state_dict"].items()}
    x = fx["x"].clone().requires_grad_(True)
    y = O.mlp(x, sd["fc1.weight"], sd["fc1.bias"], sd["fc2.weight"], sd["fc2.bias"], act=name)
    torch.testing.assert_close(y, fx["y"], atol=2e-6, rtol=1e-5)
    (y * fx["wgt"]).sum().backward()
    torch.testing.assert_close(x.grad, fx["dx"], atol=1e-5, rtol=1e-4)
    for k, g in fx["grads"].items():
        torch.testing.assert_close(sd[k].grad, g, atol=1e-4, rtol=1e-4, msg=lambda m: f"{k}: {m}")


def test_prefetch_normalisation_oracle_matches_reference_bit_exactly():
    """oracle normalize_u8 == timm's PrefetchLoader output (tests/golden/prefetch.pt), bit for bit; and the bf16 rounding of
    it equals the per-channel lookup table the CUDA patch gather builds on the host (same arithmetic in numpy)."""
    import numpy as np
    fx = torch.load(os.path.join(GOLDEN, "prefetch.pt"))
    got = O.normalize_u8(fx["u8"], fx["mean"], fx["std"])
    assert got.dtype == torch.float32 and torch.equal(got, fx["normalized"])
    for c, (m, s) in enumerate(zip(fx["mean"], fx["std"])):
        f = ((np.arange(256, dtype=np.float32) - np.float32(m * 255)) / np.float32(s * 255)).astype(np.float32)
        u = f.view(np.uint32).astype(np.uint64)
        lut = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)          # round to nearest even, as csrc/elementwise.cu
        want = fx["normalized"][:, c].to(torch.bfloat16).view(torch.int16).numpy().astype(np.uint16)
        assert (lut[fx["u8"][:, c].numpy()] == want).all()
