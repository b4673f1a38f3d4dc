#!/bin/bash
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "folded" 2>&1 | tail -15 ) > gpurun_out/r2l_fold.log 2>&1
( timeout 300 python tools/ln_fold_gemm_bench.py ) > gpurun_out/r2l_gemm.log 2>&1
( timeout 600 python tools/ln_fold_ab.py ) > gpurun_out/r2l_fold_ab.log 2>&1
tail -5 gpurun_out/r2l_fold.log; cat gpurun_out/r2l_gemm.log gpurun_out/r2l_fold_ab.log
