"""cProfile of the host side of training steps (where does the enqueue time go?)."""
import cProfile
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-image-models_b200"))
from b200timm.vision_transformer import create_model  # noqa: E402

model = create_model("vit_base_patch16_224").cuda().to(torch.bfloat16).train()
opt = torch.optim.SGD(model.parameters(), lr=1e-4, momentum=0.9)
x = torch.randn(256, 3, 224, 224, device="cuda").bfloat16()
y = torch.empty(256, dtype=torch.long, device="cuda").random_(1000)


def step():
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(model(x).float(), y)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
