#!/bin/bash
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_plugin.py -m gpu -x -q > gpurun_out/dbg.log 2>&1; tail -8 gpurun_out/dbg.log
for f in 1 0; do B200_IM2COL_FREE=$f timeout 300 python bench.py --mode infer --no-cpu-baseline --no-script-records 2>&1 | tail -1 | cut -c1-400; done
