#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_optim.py -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -12
timeout 300 python tools/optim_bench.py --dtype bf16 2>&1 | tail -2
timeout 300 python tools/optim_bench.py --dtype fp32 2>&1 | tail -2
