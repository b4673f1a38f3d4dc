#!/bin/bash
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2o_bench.log 2>&1
( timeout 120 python __graft_entry__.py smoke ) > gpurun_out/r2o_smoke.log 2>&1
tail -2 gpurun_out/r2o_smoke.log; grep -E "^real" gpurun_out/r2o_bench.log; grep '^{' gpurun_out/r2o_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','launches_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['whole_step_frac'])
print('infer', d['infer']['value'], d['infer']['e2e']['value'], d['infer']['roofline']['frac'], d['infer']['roofline']['whole_step_frac'], d['infer']['roofline'].get('traffic'))
print(d.get('cpu_baseline')); print(d.get('plugin_benchmark_py')); print(d.get('other_configs')); print(d['clocks'])
"
