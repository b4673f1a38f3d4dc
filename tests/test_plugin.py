"""The timm plugin (class-level forward rebinding, SURVEY.md 8b seam 3) on the CPU: rebinding / restoring, and that CPU
and fp32 calls keep the reference's own forward.  timm comes from oracle/_ref (the unmodified reference installed by
oracle/install_ref.py, which also ships to the GPU box; tests/test_gpu_plugin.py is the on-hardware half)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import install_ref  # noqa: E402

if install_ref.install(quiet=True) is None or not install_ref.add_to_path():
    pytest.skip("neither oracle/_ref nor /root/reference is available", allow_module_level=True)
import timm  # noqa: E402


@pytest.fixture()
def plugin():
    from b200timm import plugin as P
    P.install()
    yield P
    P.uninstall()


def test_install_rebinds_and_uninstall_restores(plugin):
    from timm.layers import Attention, AttentionRope, LayerNorm, Mlp, RmsNorm
    from timm.models.vision_transformer import Block, VisionTransformer
    for cls, attr in ((Attention, "forward"), (AttentionRope, "forward"), (Mlp, "forward"), (LayerNorm, "forward"),
                      (RmsNorm, "forward"),
                      (Block, "forward"),
                      (VisionTransformer, "forward_features"), (VisionTransformer, "forward")):
        assert hasattr(getattr(cls, attr), "_b200_original"), f"{cls.__name__}.{attr} not rebound"
    plugin.uninstall()
    assert not hasattr(Block.forward, "_b200_original")
    plugin.install()


def test_cpu_and_fp32_calls_keep_the_reference_path(plugin):
    """BASELINE configs[0] (vit_tiny, fp32, CPU, benchmark.py plumbing) must be untouched by the plugin."""
    torch.manual_seed(0)
    model = timm.create_model("vit_tiny_patch16_224", num_classes=10).eval()
    x = torch.randn(1, 3, 224, 224)
    with torch.no_grad():
        y_plugin = model(x)
    plugin.uninstall()
    with torch.no_grad():
        y_ref = model(x)
    plugin.install()
    assert torch.equal(y_plugin, y_ref)
    # parameters and state_dict keys are timm's own: nothing was replaced or re-parented
    assert list(model.state_dict())[:4] == ["cls_token", "pos_embed", "patch_embed.proj.weight", "patch_embed.proj.bias"]


def test_training_and_grad_calls_defer_to_timm_until_backward_is_registered(plugin):
    model = timm.create_model("vit_tiny_patch16_224", num_classes=4, img_size=32).train()
    out = model(torch.randn(2, 3, 32, 32))
    out.mean().backward()
    assert all(p.grad is not None for p in model.parameters())


def test_mirror_and_timm_state_dicts_are_interchangeable():
    from b200timm.vision_transformer import create_model
    ours = create_model("vit_tiny_patch16_224")
    theirs = timm.create_model("vit_tiny_patch16_224")
    assert list(ours.state_dict()) == list(theirs.state_dict())
    theirs.load_state_dict(ours.state_dict(), strict=True)
    ours.load_state_dict(theirs.state_dict(), strict=True)


def test_convnext_is_rebound_and_covered_configs_are_recognised(plugin):
    """timm's ConvNeXt / ConvNeXtBlock dispatch through the plugin; the duck-typed coverage checks accept timm's
    convnext_tiny modules (patch stem, LayerNorm, Linear MLP, exact GELU) and ConvNeXt-V2 atto's (LayerNorm2d + 1x1-conv MLP
    + GRN), reject dilated blocks (output_stride=16); a CPU call keeps timm's own forward."""
    from timm.models.convnext import ConvNeXt, ConvNeXtBlock
    from b200timm import convnext as CN
    from b200timm._lib import B200Error
    assert hasattr(ConvNeXt.forward, "_b200_original") and hasattr(ConvNeXtBlock.forward, "_b200_original")
    model = timm.create_model("convnext_tiny", num_classes=5).eval()
    for stage in model.stages:
        for blk in stage.blocks:
            CN._check_block(blk)
    v2 = timm.create_model("convnextv2_atto", num_classes=5)
    CN._check_block(v2.stages[0].blocks[0])
    dilated = timm.create_model("convnext_tiny", num_classes=5, output_stride=16)
    with pytest.raises(B200Error):
        CN._check_block(dilated.stages[3].blocks[1])
    x = torch.randn(1, 3, 64, 64)
    with torch.no_grad():
        y_plugin = model(x)
    plugin.uninstall()
    with torch.no_grad():
        y_ref = model(x)
    plugin.install()
    assert torch.equal(y_plugin, y_ref)
    ours = CN.create_model("convnext_tiny", num_classes=5)
    assert list(ours.state_dict()) == list(model.state_dict())


def test_convert_to_b200_rebinds_one_model_only():
    """Seam 2: instance-level conversion leaves classes and other models alone and keeps CPU calls on timm's path."""
    from b200timm import plugin as P
    P.uninstall()
    from timm.models.vision_transformer import Block
    a = timm.create_model("vit_tiny_patch16_224", num_classes=3, img_size=32).eval()
    b = timm.create_model("vit_tiny_patch16_224", num_classes=3, img_size=32).eval()
    keys = list(a.state_dict())
    P.convert_to_b200(a)
    assert hasattr(a.blocks[0].forward, "_b200_original") and hasattr(a.blocks[0].attn.forward, "_b200_original")
    assert not hasattr(b.blocks[0].forward, "_b200_original") and not hasattr(Block.forward, "_b200_original")
    assert list(a.state_dict()) == keys
    x = torch.randn(2, 3, 32, 32)
    b.load_state_dict(a.state_dict())
    with torch.no_grad():
        assert torch.equal(a(x), b(x))          # CPU fp32 input: the dispatcher defers to timm's forward


def test_optimizers_register_with_timms_factory():
    """register_with_timm() adds the libb200timm optimizers to timm's registry, so create_optimizer_v2 / train.py --opt
    build them with timm's own param-group logic (no weight decay on 1-d parameters, _param_groups.py:19-70)."""
    import torch
    from b200timm import optim as BO
    from timm.optim import create_optimizer_v2, list_optimizers
    BO.register_with_timm()
    assert {"b200sgd", "b200momentum", "b200adamw", "b200adam"} <= set(list_optimizers())
    m = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.LayerNorm(8))
    o = create_optimizer_v2(m, "b200sgd", lr=0.1, weight_decay=1e-4, momentum=0.9)
    assert isinstance(o, BO.SGD) and o.defaults["nesterov"] is True
    assert sorted(g["weight_decay"] for g in o.param_groups) == [0.0, 1e-4]
    o = create_optimizer_v2(m, "b200adamw", lr=1e-3, weight_decay=0.05)
    assert isinstance(o, BO.AdamW) and o.param_groups[0]["betas"] == (0.9, 0.999)


def test_model_ema_rebinding_keeps_cpu_models_on_timms_path():
    """After install(), ModelEmaV3.apply_update_ is the dispatcher; a CPU model still takes timm's own foreach lerp
    (the reference path for CPU tensors) and produces the reference result."""
    import torch
    from b200timm import plugin as P
    from timm.utils import ModelEmaV3
    P.install()
    assert hasattr(ModelEmaV3.apply_update_, "_b200_original")
    m = torch.nn.Linear(4, 3)
    ema = ModelEmaV3(m, decay=0.9)
    before = {k: v.clone() for k, v in ema.module.state_dict().items()}
    with torch.no_grad():
        m.weight.add_(1.0)
    ema.update(m)
    torch.testing.assert_close(ema.module.weight, before["weight"] + 0.1, atol=1e-6, rtol=1e-6)
    P.uninstall()
    assert not hasattr(ModelEmaV3.apply_update_, "_b200_original")
    P.install()


def test_variant_modules_are_rebound_and_cpu_calls_stay_on_timm(plugin):
    """Every variant class the table gained late in round 1 (gated MLPs, attention-pool head, qk-norm attention, GRN
    ConvNeXt blocks): rebound, and a CPU / fp32 call still produces timm's own result bit for bit."""
    from timm.layers import Attention, AttentionPoolLatent, GluMlp, LayerNorm, SwiGLU, SwiGLUPacked
    from timm.models.convnext import ConvNeXt
    torch.manual_seed(0)
    mods = [(SwiGLU(32, 64), torch.randn(2, 5, 32)), (SwiGLUPacked(32, 128), torch.randn(2, 5, 32)),
            (GluMlp(32, 128), torch.randn(2, 5, 32)), (AttentionPoolLatent(64, num_heads=1, norm_layer=LayerNorm), torch.randn(2, 9, 64)),
            (Attention(64, num_heads=1, qk_norm=True, scale_norm=True, norm_layer=LayerNorm), torch.randn(2, 9, 64)),
            (ConvNeXt(depths=(1, 1, 1, 1), dims=(8, 16, 24, 32), num_classes=5, use_grn=True, ls_init_value=None), torch.randn(2, 3, 32, 32)),
            (timm.create_model("vit_tiny_patch16_224", img_size=32, num_classes=4, global_pool="map", class_token=False,
                               act_layer="gelu_tanh", qk_norm=True), torch.randn(2, 3, 32, 32))]
    for cls in (SwiGLU, GluMlp, AttentionPoolLatent, Attention):
        assert hasattr(cls.forward, "_b200_original"), cls.__name__
    outs = []
    with torch.no_grad():
        for m, x in mods:
            outs.append(m.eval()(x))
    plugin.uninstall()
    with torch.no_grad():
        for (m, x), y in zip(mods, outs):
            assert torch.equal(m(x), y), type(m).__name__
    plugin.install()


def test_benchmark_py_cpu_plumbing_through_the_launcher(tmp_path):
    """BASELINE.json configs[0]: `benchmark.py --model vit_tiny_patch16_224 --bench inference -b 1 --device cpu --precision
    float32` (unmodified script) under `python -m b200timm.run`: the plugin is installed and every call passes through
    to timm's forward (no kernel launch, no deferral)."""
    import json
    report, res = str(tmp_path / "plugin.json"), str(tmp_path / "res.json")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([install_ref.TARGET, os.path.join(ROOT, "pytorch-image-models_b200")]),
               B200TIMM_REPORT=report, B200TIMM_STRICT="1", OMP_NUM_THREADS="4")
    cmd = [sys.executable, "-m", "b200timm.run", os.path.join(install_ref.TARGET, "scripts", "benchmark.py"), "--model",
           "vit_tiny_patch16_224", "--bench", "inference", "-b", "1", "--device", "cpu", "--precision", "float32",
           "--num-warm-iter", "1", "--num-bench-iter", "5", "--results-file", res, "--results-format", "json"]
    r = subprocess.run(cmd, env=env, cwd="/tmp", stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    out = json.load(open(res))
    out = out[0] if isinstance(out, list) else out
    assert out["infer_samples_per_sec"] > 0
    st = json.load(open(report))
    assert st["kernel_launches"] == 0 and not st["fast"] and not st["deferred"]
    assert st["passthrough"].get("timm.models.vision_transformer.VisionTransformer.forward", 0) >= 6


def test_deferrals_are_counted_and_warned_once(plugin, monkeypatch):
    """A covered-dtype call the kernels reject is deferred to timm WITH a warning and shows up in stats(); strict mode
    raises.  (Driven on the CPU by faking the `_fast` gate: the fast function then fails on the CPU tensor.)"""
    import warnings
    from b200timm._lib import B200Error
    monkeypatch.setattr(plugin, "_fast", lambda x: True)
    plugin.reset_stats()
    m = timm.layers.Mlp(8, 16).eval()
    x = torch.randn(2, 3, 8)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        with torch.no_grad():
            y1 = m(x)
            y2 = m(x)
    assert torch.equal(y1, y2)
    d = plugin.stats()["deferred"]
    assert sum(d["timm.layers.Mlp.forward"].values()) == 2
    assert len([x for x in w if "deferred to timm" in str(x.message)]) == 1
    monkeypatch.setattr(plugin, "_STRICT", True)
    with pytest.raises(B200Error), torch.no_grad():
        m(x)
