#!/bin/bash
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( timeout 600 python -m pytest tests/test_gpu_convnext.py tests/test_gpu_kernels.py -q -k "convnext or dwconv or folded or wgrad" 2>&1 | tail -30 ) > gpurun_out/r2n_cnx_tests.log 2>&1
( timeout 300 python tools/bench_model.py convnext_base --batch 256 --detail ) > gpurun_out/r2n_convnext_fold.log 2>&1
( B200_LN_FOLD=0 timeout 300 python tools/bench_model.py convnext_base --batch 256 ) > gpurun_out/r2n_convnext_nofold.log 2>&1
( timeout 300 python tools/bench_model.py convnext_base --batch 256 ) > gpurun_out/r2n_convnext_fold2.log 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/r2n_tests.log 2>&1
tail -8 gpurun_out/r2n_cnx_tests.log; tail -22 gpurun_out/r2n_convnext_fold.log; tail -1 gpurun_out/r2n_convnext_nofold.log; tail -1 gpurun_out/r2n_convnext_fold2.log; tail -5 gpurun_out/r2n_tests.log
