#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_glu.py tests/test_gpu_data.py tests/test_gpu_optim.py tests/test_gpu_loss.py tests/test_gpu_naflex.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -25
