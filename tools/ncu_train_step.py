"""ONE ViT-B/16 training step (batch 256, bf16: zero_grad + forward + cross entropy + backward + Nesterov SGD, the step
bench.py times) -- or, with --infer, one inference forward -- between cudaProfilerStart / Stop, eagerly launched, for
`ncu --profile-from-start off` (tools/gpu_ncu_step.sh):
    ncu --metrics gpu__time_duration.sum ... python tools/ncu_train_step.py              # launch list of a step
    ncu --set full ...                       python tools/ncu_train_step.py --depth 1    # every kernel type once
A number printed under ncu is never a bench value."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-image-models_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=12)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--infer", action="store_true")
    args = ap.parse_args()
    from b200timm import loss as BL, optim as BO
    from b200timm.train_step import TrainStep
    from b200timm.vision_transformer import create_model
    torch.manual_seed(42)
    model = create_model("vit_base_patch16_224", depth=args.depth).cuda().to(torch.bfloat16)
    xs = [torch.randn(args.batch, 3, 224, 224, device="cuda").bfloat16() for _ in range(2)]
    ys = [torch.randint(0, 1000, (args.batch,), device="cuda") for _ in range(2)]
    if args.infer:
        model.eval()
        with torch.no_grad():
            for i in range(3):
                model(xs[i & 1])
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStart()
            model(xs[1])
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
        return
    model.train()
    opt = BO.SGD(model.parameters(), lr=1e-4, momentum=0.9, nesterov=True)
    step = TrainStep(model, BL.CrossEntropyLoss(), opt, xs[0], ys[0], graph=False)
    for i in range(3):
        step(xs[i & 1], ys[i & 1])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    step(xs[1], ys[1])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
