#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_naflex.py -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -8
timeout 600 python tools/bench_model.py naflex:vit_large_patch14_clip_336 --batch 64 --steps 5 --warmup 2 --train 2>&1 | tail -2
