"""One launch of each kernel added late in round 2, between cudaProfilerStart / Stop, for `ncu --set full --profile-from-start off`
(tools/gpu_ncu_r02b.sh): the 128-wide attention forward / backward (ViT-H/14 geometry), the depthwise convolution with LayerNorm
statistics and its consumer GEMM (ConvNeXt-B stage 3), ViT-L's qkv weight gradient as row chunks."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-image-models_b200"))
from b200timm import ops  # noqa: E402

B, N, H, hd = 64, 257, 16, 128
qkv = torch.randn(B, N, 3 * H * hd, device="cuda").bfloat16()
dout = torch.randn(B, N, H * hd, device="cuda").bfloat16()
x = torch.randn(256, 14, 14, 512, device="cuda").bfloat16()
wt = (torch.randn(7, 7, 512, device="cuda") * 0.2).bfloat16()
bias = torch.randn(512, device="cuda").bfloat16()
w1 = (torch.randn(2048, 512, device="cuda") * 0.05).bfloat16()
c1, c2 = torch.randn(2048, device="cuda"), torch.randn(2048, device="cuda")


def run():
    out, lse = ops.attention_fwd(qkv, B, N, H, hd, hd ** -0.5, need_lse=True)
    ops.attention_bwd(qkv, out, dout, lse, B, N, H, hd, hd ** -0.5)
    y, st = ops.dwconv_nhwc_ln(x, wt, bias)
    ops.linear_fwd(y.view(-1, 512), w1, c2, epilogue=ops.EPI_BIAS_GELU, ln_in=(st, c1, 1e-6))


run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
