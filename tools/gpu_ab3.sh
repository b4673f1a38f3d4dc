#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_acts.py tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -4
one() { python bench.py $2 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], d.get('roofline',{}).get('frac'))"; }
for i in 1 2 3; do
  B200TIMM_LIB=/root/repo/ab/libb200timm_prev.so one "prev infer" "--steps 20 --warmup 5 --no-cpu-baseline"
  one "new  infer" "--steps 20 --warmup 5 --no-cpu-baseline"
done
