#!/bin/bash
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_models.py -q -x -k "im2col" 2>&1 | tail -25 > gpurun_out/r2o.log
timeout 900 python -m pytest tests/test_gpu_plugin.py -q -x -k "variant_models or convnext" 2>&1 | tail -12 >> gpurun_out/r2o.log
for f in 1 0; do B200_IM2COL_FREE=$f timeout 200 python bench.py --mode infer --steps 20 --warmup 5 --no-cpu-baseline --no-script-records 2>&1 | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('im2col_free=$f', d['value'], d['ms_per_step'], d['launches_per_step'])" >> gpurun_out/r2o.log; done
cat gpurun_out/r2o.log
