#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train.py tests/test_gpu_rope.py -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -4
timeout 300 python tools/train_detail.py 2>&1 | head -4
timeout 600 python bench.py --mode train --steps 10 --warmup 3 2>&1 | tail -1 | cut -c80-220
