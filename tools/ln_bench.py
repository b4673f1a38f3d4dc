"""LayerNorm forward / backward timing at the ViT-B training shape ([50432, 768] bf16), CUDA events, L2 flushed by size."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-image-models_b200"))
from b200timm import ops  # noqa: E402


def main():
    for rows, cols in ((50432, 768), (36928, 1024)):
        xs = [torch.randn(rows, cols, device="cuda").bfloat16() for _ in range(3)]
        dys = [torch.randn(rows, cols, device="cuda").bfloat16() for _ in range(3)]
        res = [torch.randn(rows, cols, device="cuda").bfloat16() for _ in range(3)]
        w = torch.randn(cols, device="cuda").bfloat16()
        b = torch.randn(cols, device="cuda").bfloat16()
        y, mean, rstd = ops.layernorm_fwd(xs[0], w, b, 1e-6, save_stats=True)

        def timeit(fn, n=30):
            for i in range(5):
                fn(i)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for i in range(n):
                fn(i)
            t1.record()
            torch.cuda.synchronize()
            return t0.elapsed_time(t1) / n * 1e3
        unit = rows * cols * 2 / 1e6
        f = timeit(lambda i: ops.layernorm_fwd(xs[i % 3], w, b, 1e-6, save_stats=True))
        bw = timeit(lambda i: ops.layernorm_bwd(dys[i % 3], xs[i % 3], w, mean, rstd, dres=res[i % 3]))
        bw0 = timeit(lambda i: ops.layernorm_bwd(dys[i % 3], xs[i % 3], w, mean, rstd))
        print(f"[{rows}, {cols}] ring={os.environ.get('B200_LN_BWD_RING', '1')}: fwd {f:.1f} us ({2 * unit / f / 1e3:.2f} TB/s), "
              f"bwd+res {bw:.1f} us ({4 * unit / bw / 1e3:.2f} TB/s), bwd {bw0:.1f} us ({3 * unit / bw0 / 1e3:.2f} TB/s)")


if __name__ == "__main__":
    main()
