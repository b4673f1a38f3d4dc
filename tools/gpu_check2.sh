#!/bin/bash
# wide-head attention tests first (short timeout: new kernels), then the whole GPU suite without -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( time timeout 600 python -m pytest tests/test_gpu_wide_heads.py -q 2>&1 | tail -60 ) > gpurun_out/r2i_wide.log 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_wide_heads.py 2>&1 | tail -80 ) > gpurun_out/r2i_tests.log 2>&1
tail -25 gpurun_out/r2i_wide.log; tail -15 gpurun_out/r2i_tests.log
