#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train.py -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tools/microbench.py 2>&1 | grep -E "gelu|qkv"
timeout 300 python tools/train_detail.py 2>&1 | head -8
