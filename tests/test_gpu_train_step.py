"""b200timm.train_step.TrainStep (the whole training step as one CUDA graph) and the weight-cache invalidation after
raw-pointer parameter writes (optimizer / EMA kernels, graph replays)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _small_vit(num_classes=16, **kw):
    from b200timm.vision_transformer import VisionTransformer
    torch.manual_seed(0)
    cfg = dict(img_size=64, patch_size=16, embed_dim=128, depth=2, num_heads=2, num_classes=num_classes)
    cfg.update(kw)
    return VisionTransformer(**cfg).cuda().to(torch.bfloat16)


def test_graphed_step_matches_eager_steps():
    """Same initial weights, same batches: losses of N graphed steps == losses of N eager steps (the wgrad split-K sums
    are fp32 reduce-adds whose order is not fixed, hence a tolerance instead of bit equality), parameters stay close.
    The graphed TrainStep runs 3 eager warm-up steps on its first batch before capturing; the eager run does the same."""
    from b200timm import loss as BL, optim as BO
    from b200timm.train_step import TrainStep
    xs = [torch.randn(8, 3, 64, 64, device="cuda").bfloat16() for _ in range(4)]
    ys = [torch.randint(0, 16, (8,), device="cuda") for _ in range(4)]
    ref = _small_vit().train()
    gm = copy.deepcopy(ref)
    opt = BO.SGD(gm.parameters(), lr=1e-2, momentum=0.9, nesterov=True)
    step = TrainStep(gm, BL.CrossEntropyLoss(), opt, xs[0], ys[0], graph=True, warmup=3)
    graphed = [float(step(xs[i % 4], ys[i % 4])) for i in range(6)]
    opt = BO.SGD(ref.parameters(), lr=1e-2, momentum=0.9, nesterov=True)
    st = TrainStep(ref, BL.CrossEntropyLoss(), opt, xs[0], ys[0], graph=False)
    for _ in range(3):
        st(xs[0], ys[0])
    eager = [float(st(xs[i % 4], ys[i % 4])) for i in range(6)]
    torch.testing.assert_close(torch.tensor(graphed), torch.tensor(eager), atol=2e-2, rtol=2e-2)
    for (n, a), (_, b) in zip(gm.named_parameters(), ref.named_parameters()):
        torch.testing.assert_close(a.float(), b.float(), atol=3e-2, rtol=3e-2, msg=lambda m: f"{n}: {m}")
    assert step.launches > 50


def test_capture_on_resident_buffers_and_launch_count():
    from b200timm import loss as BL, optim as BO
    from b200timm.train_step import TrainStep
    model = _small_vit().train()
    opt = BO.SGD(model.parameters(), lr=1e-3, momentum=0.9, nesterov=True)
    xs = [torch.randn(4, 3, 64, 64, device="cuda").bfloat16() for _ in range(2)]
    ys = [torch.randint(0, 16, (4,), device="cuda") for _ in range(2)]
    step = TrainStep(model, BL.CrossEntropyLoss(), opt, xs[0], ys[0])
    runners = [step.capture_on(xs[i], ys[i]) for i in range(2)]
    before = [p.detach().clone() for p in model.parameters()]
    l0 = float(runners[0]())
    l1 = float(runners[1]())
    assert l0 == l0 and l1 == l1 and l0 != l1
    assert any(not torch.equal(a, p.detach()) for a, p in zip(before, model.parameters()))
    assert step.launches > 50 and not step.lr_changed()
    opt.param_groups[0]["lr"] = 5e-4
    assert step.lr_changed()


@pytest.mark.parametrize("kind", ["vit10", "convnext", "vit_amp"])
def test_eval_after_optimizer_step_sees_the_new_weights(kind):
    """derived weight copies (bf16 casts of fp32 masters, padded 10-class head, transposed depthwise /
    conv weights) are cached by the inference path; the optimizer / EMA kernels write parameters through raw pointers.
    eval -> step -> eval must equal a forward with every cache dropped."""
    from b200timm import loss as BL, optim as BO
    torch.manual_seed(0)
    if kind == "convnext":
        from b200timm.convnext import ConvNeXt
        model = ConvNeXt(depths=(1, 1, 1, 1), dims=(64, 128, 192, 256), num_classes=10).cuda().to(torch.bfloat16)
        x = torch.randn(4, 3, 64, 64, device="cuda").bfloat16()
    else:
        model = _small_vit(num_classes=10)
        if kind == "vit_amp":
            model = model.float()
        x = torch.randn(4, 3, 64, 64, device="cuda").bfloat16()
    y = torch.randint(0, 10, (4,), device="cuda")
    opt = BO.SGD(model.parameters(), lr=0.5, momentum=0.9, nesterov=True)

    def evaluate():
        model.eval()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=kind == "vit_amp"):
            return model(x if kind != "vit_amp" else x.float()).float().clone()

    def drop_caches():
        for m in model.modules():
            for k in [k for k in vars(m) if k.startswith("_b200_")]:
                delattr(m, k)

    e0 = evaluate()
    model.train()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=kind == "vit_amp"):
        out = model(x if kind != "vit_amp" else x.float())
    BL.CrossEntropyLoss()(out, y).backward()
    opt.step()
    e1 = evaluate()
    drop_caches()
    from b200timm import ops
    ops.note_params_written()            # forces every cache key to miss
    e2 = evaluate()
    assert not torch.equal(e0, e1), "the optimizer step did not change the evaluation output"
    assert torch.equal(e1, e2), f"stale cached weights: max diff {(e1 - e2).abs().max().item():.4g}"
