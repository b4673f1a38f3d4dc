#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q --no-header -p no:cacheprovider -k layernorm 2>&1 | tail -3
for m in 0 1 0 1; do echo "== two_rows $m"; B200_LN_TWO_ROWS=$m timeout 300 python tools/microbench.py 2>&1 | grep "layernorm fwd" | cut -c1-120; B200_LN_TWO_ROWS=$m timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c80-200; done
