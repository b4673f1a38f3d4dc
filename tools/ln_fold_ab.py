"""Same-process A/B of the folded LayerNorm (functional._LN_FOLD) on the bench workload: ViT-B/16 bf16 inference, batch 256,
CUDA-graph replays, alternating rounds; plus the per-GEMM event times of one eager forward in each mode."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-image-models_b200"))
from b200timm import functional as BF, ops  # noqa: E402
from b200timm.vision_transformer import create_model  # noqa: E402

torch.manual_seed(0)
model = create_model("vit_base_patch16_224").cuda().to(torch.bfloat16).eval()
img = torch.randn(256, 3, 224, 224, device="cuda").bfloat16()
graphs = {}
for fold in (True, False):
    BF._LN_FOLD = fold
    with torch.no_grad():
        for _ in range(3):
            model(img)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = model(img)
    graphs[fold] = (g, out)
BF._LN_FOLD = True
big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(g, n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for rnd in range(4):
    a, b = run(graphs[True][0]), run(graphs[False][0])
    print(f"round {rnd}: folded {a:.3f} ms ({256 / a * 1e3:.0f} img/s)   LayerNorm kernels {b:.3f} ms ({256 / b * 1e3:.0f} img/s)   ratio {b / a:.4f}")
d = (graphs[True][1].float() - graphs[False][1].float()).abs().max().item()
print("max |logit difference| folded vs unfolded:", d, " logits absmax", graphs[False][1].float().abs().max().item())
# per-GEMM times (CUDA events around every b200_gemm of one eager forward)
for fold in (True, False):
    BF._LN_FOLD = fold
    ops._gemm_timing = []
    with torch.no_grad():
        model(img)
    torch.cuda.synchronize()
    t = [(e0.elapsed_time(e1), fl) for e0, e1, fl in ops._gemm_timing]
    ops._gemm_timing = None
    blk = t[1:5] if len(t) > 5 else t
    mid = t[1 + 4 * 6: 1 + 4 * 6 + 4]
    print("fold" if fold else "plain", "block 6 GEMMs (qkv, proj, fc1, fc2) us:", [round(x[0] * 1e3, 1) for x in mid],
          " all GEMMs ms:", round(sum(x[0] for x in t), 3))
BF._LN_FOLD = True
