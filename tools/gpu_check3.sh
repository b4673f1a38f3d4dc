#!/bin/bash
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( time timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "folded" 2>&1 | tail -40 ) > gpurun_out/r2j_fold.log 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 ) > gpurun_out/r2j_tests.log 2>&1
( timeout 300 python tools/wide_head_bench.py ) > gpurun_out/r2j_wide_bench.log 2>&1
( time B200_LN_FOLD=0 timeout 600 python bench.py --steps 20 --warmup 5 --mode infer ) > gpurun_out/r2j_bench_nofold.log 2>&1
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2j_bench.log 2>&1
tail -12 gpurun_out/r2j_fold.log; tail -8 gpurun_out/r2j_tests.log; cat gpurun_out/r2j_wide_bench.log
for f in gpurun_out/r2j_bench_nofold.log gpurun_out/r2j_bench.log; do grep '^{' $f | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','launches_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('whole_step_frac'))
i=d.get('infer')
if i: print('infer', i['value'], i['e2e']['value'], i['roofline']['frac'], i['roofline']['whole_step_frac'], i.get('launches_per_step'))
print(d.get('other_configs')); print(d['clocks'])
"; done
