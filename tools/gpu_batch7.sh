#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_acts.py tests/test_gpu_models.py tests/test_gpu_convnext.py tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -25
timeout 300 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('infer', d['value'], d['ms_per_step'], d['roofline']['frac'])"
timeout 300 python bench.py --mode train --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('train', d['value'], d['ms_per_step'])"
