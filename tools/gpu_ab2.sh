#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
one() { python bench.py $2 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], d.get('roofline',{}).get('frac'))"; }
for i in 1 2; do
  B200TIMM_LIB=/root/repo/ab/libb200timm_prev.so one "prev infer" "--steps 20 --warmup 5 --no-cpu-baseline"
  one "new  infer" "--steps 20 --warmup 5 --no-cpu-baseline"
done
B200TIMM_LIB=/root/repo/ab/libb200timm_prev.so one "prev train" "--mode train --steps 10 --warmup 3"
one "new  train" "--mode train --steps 10 --warmup 3"
