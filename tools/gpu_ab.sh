#!/bin/bash
# same-box A/B of the GEMM CTA-pair mode
cd /root/repo; mkdir -p gpurun_out
for mode in 0 1 0 1; do
  echo "== B200_GEMM_2CTA=$mode"
  B200_GEMM_2CTA=$mode timeout 300 python tools/microbench.py 2>&1 | grep gemm | cut -c1-120
  B200_GEMM_2CTA=$mode timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c80-200
done
B200_GEMM_2CTA=0 timeout 600 python bench.py --mode train --steps 10 --warmup 3 2>&1 | tail -1 | cut -c80-200
B200_GEMM_2CTA=1 timeout 600 python bench.py --mode train --steps 10 --warmup 3 2>&1 | tail -1 | cut -c80-200
