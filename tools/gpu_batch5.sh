#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_convnext.py tests/test_gpu_qknorm.py tests/test_gpu_data.py tests/test_gpu_train.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -25
