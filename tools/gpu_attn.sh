#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train.py tests/test_gpu_naflex.py -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tools/train_detail.py 2>&1 | head -12
timeout 600 python bench.py --mode train --steps 10 --warmup 3 2>&1 | tail -1 | cut -c80-220
