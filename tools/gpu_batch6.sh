#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_models.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -25
timeout 300 python tools/bench_model.py convnextv2_base --batch 256 --steps 5 --warmup 2 2>&1 | tail -1
timeout 300 python tools/bench_model.py convnextv2_base --batch 128 --steps 5 --warmup 2 --train 2>&1 | tail -1
