#!/bin/bash
# same-box A/B of the GEMM modes: 2 = CTA pairs, 32-column rounds / 6 stages; 1 = CTA pairs, 64-column rounds / 5 stages
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train.py -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -5
for mode in 2 1 2 1; do
  echo "== B200_GEMM_2CTA=$mode"
  B200_GEMM_2CTA=$mode timeout 300 python tools/microbench.py 2>&1 | grep gemm | cut -c1-120
  B200_GEMM_2CTA=$mode timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c80-200
done
B200_GEMM_2CTA=2 timeout 600 python bench.py --mode train --steps 10 --warmup 3 2>&1 | tail -1 | cut -c80-200
B200_GEMM_2CTA=1 timeout 600 python bench.py --mode train --steps 10 --warmup 3 2>&1 | tail -1 | cut -c80-200
