"""Optimizer step on the GPU: b200timm.optim.{SGD, AdamW, Adam} and clip_grad_norm_ through the C ABI
against (a) the reference's own trajectories (tests/golden/optim.pt: timm create_optimizer_v2 + dispatch_clip_grad, fp32)
and (b) the CPU oracle for bf16 parameters with fp32 state.

Tolerances: fp32 parameters -- atol 3e-7, rtol 3e-6 after three steps (fused-multiply-add contraction is the only
difference from torch's kernel chain).  bf16 parameters -- at most one bf16 ulp on at most 0.5 % of the elements (an fp32
rounding difference that straddles a bf16 rounding boundary), everything else bit-exact."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _make(case, params):
    from b200timm import optim as BO
    kw = dict(case["kwargs"])
    if case["opt"] in ("sgd", "momentum"):
        return BO.SGD(params, nesterov=(case["opt"] == "sgd"), **kw)
    return (BO.AdamW if case["opt"] == "adamw" else BO.Adam)(params, **kw)


def _run(case, init, grads, dtype, fused_clip=True, offset=0):
    from b200timm import optim as BO
    if offset:      # parameters / gradients as misaligned views into flat buffers (DDP bucket views): scalar path
        params = []
        for t in init:
            flat = torch.zeros(t.numel() + offset, dtype=dtype, device="cuda")
            flat[offset:] = t.flatten().to(dtype)
            params.append(torch.nn.Parameter(flat[offset:].view(t.shape)))
    else:
        params = [torch.nn.Parameter(t.cuda().to(dtype)) for t in init]
    opt = _make(case, params)
    traj, norms = [], []
    for gk in grads:
        for p, g in zip(params, gk):
            p.grad = g.cuda().to(dtype)
        if case["clip"] is not None:
            norms.append(BO.clip_grad_norm_(params, case["clip"], optimizer=opt if fused_clip else None))
        opt.step()
        traj.append([p.detach().clone() for p in params])
    return traj, [float(n) for n in norms], opt


CASES = ["sgd_nesterov_wd", "momentum_dampening", "sgd_plain", "adamw", "adam_l2", "adamw_clip", "sgd_clip"]


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("offset", [0, 1])
def test_fp32_parameters_match_reference_trajectory(name, offset):
    fx = torch.load(os.path.join(GOLDEN, "optim.pt"))
    case = fx["cases"][name]
    traj, norms, _ = _run(case, fx["init"], fx["grads"], torch.float32, offset=offset)
    for k, (got, ref) in enumerate(zip(traj, case["params"])):
        for a, b in zip(got, ref):
            torch.testing.assert_close(a.cpu(), b, atol=3e-7, rtol=3e-6, msg=lambda m: f"{name} step {k}: {m}")
    for a, b in zip(norms, case["norms"]):
        assert abs(a - b) <= 1e-5 * b, (name, a, b)


@pytest.mark.parametrize("name", ["adamw_clip", "sgd_clip"])
def test_inplace_clip_equals_fused_clip_and_scales_gradients(name):
    from b200timm import optim as BO
    fx = torch.load(os.path.join(GOLDEN, "optim.pt"))
    case = fx["cases"][name]
    fused, _, _ = _run(case, fx["init"], fx["grads"], torch.float32, fused_clip=True)
    inplace, _, _ = _run(case, fx["init"], fx["grads"], torch.float32, fused_clip=False)
    for a, b in zip(fused[-1], inplace[-1]):
        torch.testing.assert_close(a, b, atol=3e-7, rtol=3e-6)
    # the in-place form leaves clipped gradients behind, like torch.nn.utils.clip_grad_norm_
    ps = [torch.nn.Parameter(t.cuda()) for t in fx["init"]]
    for p, g in zip(ps, fx["grads"][2]):
        p.grad = g.cuda().clone()
    ref = [p.grad.clone() for p in ps]
    total = torch.nn.utils.get_total_norm([p.grad for p in ps])
    got = BO.clip_grad_norm_(ps, 0.25)
    assert abs(float(got) - float(total)) <= 1e-5 * float(total)
    coef = min(1.0, 0.25 / (float(total) + 1e-6))
    for p, r in zip(ps, ref):
        torch.testing.assert_close(p.grad, r * coef, atol=1e-7, rtol=2e-6)


@pytest.mark.parametrize("name", CASES)
def test_bf16_parameters_fp32_state_match_oracle(name):
    from test_oracle_golden import _oracle_optim_run
    fx = torch.load(os.path.join(GOLDEN, "optim.pt"))
    case = fx["cases"][name]
    traj, _, opt = _run(case, fx["init"], fx["grads"], torch.bfloat16)
    ref, _ = _oracle_optim_run(case, fx["init"], fx["grads"], round_to=torch.bfloat16)
    for st in opt.state.values():
        for k, v in st.items():
            assert not torch.is_tensor(v) or v.dtype == torch.float32, k
    n = bad = 0
    for got, want in zip(traj[-1], ref[-1]):
        g, w = got.float().cpu(), want
        assert got.dtype == torch.bfloat16
        diff = (g - w).abs()
        ulp = torch.maximum(w.abs(), torch.tensor(1e-30)) * 2.0 ** -7      # >= one bf16 ulp of the value
        assert (diff <= ulp).all(), f"{name}: more than one bf16 ulp apart"
        n += g.numel()
        bad += int((diff > 0).sum())
    assert bad <= 0.005 * n, f"{name}: {bad} of {n} elements differ"


def test_many_tensors_param_groups_and_state_dict_round_trip():
    """> 512 tensors (several launches per group), two param groups with different hyper-parameters, parameters that
    receive their first gradient late, and Optimizer.load_state_dict (which casts state to the parameter dtype)."""
    from b200timm import optim as BO
    torch.manual_seed(0)
    shapes = [(i % 7 + 1, i % 5 + 3) for i in range(700)]
    init = [torch.randn(s) for s in shapes]
    mine = [torch.nn.Parameter(t.cuda()) for t in init]
    ref = [torch.nn.Parameter(t.cuda()) for t in init]
    groups = lambda ps: [dict(params=ps[:300], weight_decay=0.0), dict(params=ps[300:], lr=3e-3)]      # noqa: E731
    o1 = BO.AdamW(groups(mine), lr=1e-3, weight_decay=0.1)
    o2 = torch.optim.AdamW(groups(ref), lr=1e-3, weight_decay=0.1)
    for k in range(3):
        for i, (a, b) in enumerate(zip(mine, ref)):
            if k == 0 and i % 3 == 0:
                a.grad = b.grad = None          # first gradient arrives at step 1: per-tensor step counts differ
                continue
            g = torch.randn(a.shape, device="cuda")
            a.grad, b.grad = g.clone(), g.clone()
        o1.step()
        o2.step()
    for a, b in zip(mine, ref):
        torch.testing.assert_close(a, b, atol=3e-7, rtol=3e-6)
    # bf16 parameters: the loaded state arrives as bf16 and is widened again before the kernel sees it
    pb = [torch.nn.Parameter(t.cuda().bfloat16()) for t in init[:8]]
    ob = BO.SGD(pb, lr=0.1, momentum=0.9)
    for p in pb:
        p.grad = torch.randn_like(p)
    ob.step()
    sd = ob.state_dict()
    ob2 = BO.SGD(pb, lr=0.1, momentum=0.9)
    ob2.load_state_dict(sd)
    ob2.step()
    assert all(st["momentum_buffer"].dtype == torch.float32 for st in ob2.state.values())
    assert all(torch.isfinite(p).all() for p in pb)


def test_rejects_cpu_and_unsupported_flags():
    from b200timm import optim as BO
    from b200timm._lib import B200Error
    p = torch.nn.Parameter(torch.randn(4))
    p.grad = torch.randn(4)
    with pytest.raises(B200Error):
        BO.SGD([p], lr=0.1).step()
    with pytest.raises(B200Error):
        BO.AdamW([p], amsgrad=True)
    with pytest.raises(ValueError):
        BO.SGD([p], lr=0.1, nesterov=True)


def test_training_step_with_fused_optimizer_matches_torch_optimizer():
    """A ViT training step (kernels' backward) followed by b200timm.optim.SGD == the same step followed by
    torch.optim.SGD with fp32 master weights (train.py --amp layout)."""
    from b200timm import optim as BO
    from b200timm.vision_transformer import VisionTransformer
    kw = dict(img_size=64, patch_size=16, embed_dim=128, depth=2, num_heads=2, num_classes=10)
    torch.manual_seed(0)
    m1 = VisionTransformer(**kw).cuda().train()
    m2 = VisionTransformer(**kw).cuda().train()
    m2.load_state_dict(m1.state_dict())
    o1 = BO.SGD(m1.parameters(), lr=0.05, momentum=0.9, nesterov=True, weight_decay=1e-4)
    o2 = torch.optim.SGD(m2.parameters(), lr=0.05, momentum=0.9, nesterov=True, weight_decay=1e-4)
    x = torch.randn(4, 3, 64, 64, device="cuda")
    y = torch.randint(0, 10, (4,), device="cuda")
    for _ in range(2):
        for m, o, clip in ((m1, o1, BO.clip_grad_norm_), (m2, o2, torch.nn.utils.clip_grad_norm_)):
            o.zero_grad(set_to_none=True)
            torch.nn.functional.cross_entropy(m(x).float(), y).backward()
            clip(list(m.parameters()), 1.0)
            o.step()
    for (n, a), b in zip(m1.named_parameters(), m2.parameters()):
        torch.testing.assert_close(a, b, atol=2e-5, rtol=1e-4, msg=lambda s: f"{n}: {s}")


# ------------------------------------------------------------------------------------------------ weight EMA
def _ema_model():
    return torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.BatchNorm1d(64), torch.nn.Linear(64, 5))


def test_model_ema_matches_reference_golden():
    """b200timm.ema.ModelEma == timm.utils.ModelEmaV3 on the golden trajectory (warm-up schedule, float entries
    interpolated by b200_ema_update, the integer BatchNorm counter copied)."""
    from b200timm.ema import ModelEma
    fx = torch.load(os.path.join(GOLDEN, "ema.pt"))
    model = _ema_model()
    model.load_state_dict(fx["init"])
    model = model.cuda()
    ema = ModelEma(model, **fx["kwargs"])
    for step, (msd, want, d) in enumerate(zip(fx["model_states"], fx["ema_states"], fx["decays"]), start=1):
        model.load_state_dict(msd)
        assert ema.get_decay(step) == d
        ema.update(model, step=step)
        got = ema.module.state_dict()
        for k in want:
            if want[k].is_floating_point():
                torch.testing.assert_close(got[k].cpu(), want[k], atol=1e-7, rtol=2e-6, msg=lambda m: f"step {step} {k}: {m}")
            else:
                assert torch.equal(got[k].cpu(), want[k]), k


@pytest.mark.parametrize("ema_dtype", [None, torch.float32])
def test_model_ema_bf16_model_matches_oracle(ema_dtype):
    """bf16 model with a bf16 (reference behaviour) or fp32 average against the oracle's lerp rounded to the average's
    dtype.  bf16 average: within one bf16 ulp everywhere and bit-exact on >= 99 % of the elements; fp32 average: atol 1e-7
    / rtol 1e-6 after three updates (the kernel's fused multiply-add rounds once, torch's mul + add twice)."""
    from oracle import optim_oracle as OO
    from b200timm.ema import ModelEma
    torch.manual_seed(5)
    model = torch.nn.Sequential(torch.nn.Linear(64, 256), torch.nn.LayerNorm(256), torch.nn.Linear(256, 9)).cuda().bfloat16()
    ema = ModelEma(model, decay=0.99, ema_dtype=ema_dtype)
    ref = {k: v.detach().cpu().clone() for k, v in ema.module.state_dict().items()}
    for _ in range(3):
        with torch.no_grad():
            for p in model.parameters():
                p.add_(torch.randn_like(p) * 0.05)
        ema.update(model)
        OO.ema_update(ref, {k: v.detach().cpu() for k, v in model.state_dict().items()}, 0.99,
                      round_to=torch.bfloat16 if ema_dtype is None else torch.float32)
    for k, v in ema.module.state_dict().items():
        assert v.dtype == (torch.bfloat16 if ema_dtype is None else torch.float32)
        g, w = v.float().cpu(), ref[k].float()
        if ema_dtype is None:
            assert ((g - w).abs() <= w.abs().clamp_min(1e-30) * 2.0 ** -7).all(), k
            assert ((g - w).abs() > 0).float().mean() <= 0.01, k
        else:                       # an fp32 average shows every fused-multiply-add rounding difference
            torch.testing.assert_close(g, w, atol=1e-7, rtol=1e-6, msg=lambda m: f"{k}: {m}")
