#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_optim.py tests/test_gpu_loss.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -12
