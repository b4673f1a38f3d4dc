#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -4
timeout 600 python tools/bench_model.py vit_large_patch14_clip_336 --batch 64 --steps 5 --warmup 2 --train --detail 2>&1 | head -8
timeout 600 python tools/bench_model.py vit_large_patch14_clip_336 --batch 64 --steps 5 --warmup 2 --train 2>&1 | tail -1
