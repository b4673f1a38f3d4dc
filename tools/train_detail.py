"""Per-op CUDA-event time split of one ViT-B/16 training step (forward + backward through the kernel path)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-image-models_b200"))
from b200timm import ops  # noqa: E402
from b200timm.vision_transformer import create_model  # noqa: E402

B = int(os.environ.get("TB", "256"))
model = create_model("vit_base_patch16_224").cuda().to(torch.bfloat16).train()
opt = torch.optim.SGD(model.parameters(), lr=1e-4, momentum=0.9)
x = torch.randn(B, 3, 224, 224, device="cuda").bfloat16()
y = torch.empty(B, dtype=torch.long, device="cuda").random_(1000)


def step():
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(model(x).float(), y)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ops.enable_op_timing(True)
t0.record()
step()
t1.record()
det = ops.op_timing_summary()
ops.enable_op_timing(False)
tot = sum(t for _, t in det.values())
print(f"step {t0.elapsed_time(t1):.2f} ms, inside our ops {tot:.2f} ms")
for k, (n, t) in sorted(det.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"  {t:8.3f} ms {100 * t / tot:5.1f}%  n={n:3d}  {k}")
