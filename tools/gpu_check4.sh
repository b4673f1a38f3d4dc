#!/bin/bash
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( timeout 600 python tools/ln_fold_ab.py ) > gpurun_out/r2k_fold_ab.log 2>&1
( time timeout 900 python -m pytest tests/test_gpu_plugin.py -q -x 2>&1 | tail -30 ) > gpurun_out/r2k_plugin.log 2>&1
cat gpurun_out/r2k_fold_ab.log; tail -6 gpurun_out/r2k_plugin.log
