#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tools/microbench.py 2>&1 | grep "layernorm" | cut -c1-120
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c80-200
