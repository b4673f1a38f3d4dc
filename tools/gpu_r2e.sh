#!/bin/bash
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train.py tests/test_gpu_acts.py tests/test_gpu_plugin.py -q -x 2>&1 | tail -15 > gpurun_out/r2e_tests.log
timeout 300 python bench.py --mode train --steps 10 --warmup 3 --no-cpu-baseline --no-script-records > gpurun_out/r2e_train.log 2>&1
timeout 120 python tools/train_detail.py > gpurun_out/r2e_detail.log 2>&1
tail -6 gpurun_out/r2e_tests.log; grep '^{' gpurun_out/r2e_train.log | cut -c1-330; head -24 gpurun_out/r2e_detail.log
