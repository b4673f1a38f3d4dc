#!/bin/bash
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention_bwd" 2>&1 | tail -5 > gpurun_out/r2i_tests.log
timeout 600 python -m pytest tests/test_gpu_data.py tests/test_gpu_naflex.py tests/test_gpu_train.py -q -x 2>&1 | tail -8 >> gpurun_out/r2i_tests.log
timeout 120 python tools/train_detail.py > gpurun_out/r2i_detail.log 2>&1
timeout 300 python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline --no-script-records > gpurun_out/r2i_train.log 2>&1
cat gpurun_out/r2i_tests.log; head -8 gpurun_out/r2i_detail.log; grep '^{' gpurun_out/r2i_train.log | cut -c1-300
