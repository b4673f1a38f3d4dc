#!/bin/bash
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_plugin.py tests/test_gpu_kernels.py tests/test_gpu_rope.py tests/test_gpu_glu.py -q 2>&1 | tail -150 > gpurun_out/r2g_tests.log
tail -5 gpurun_out/r2g_tests.log
grep -E "^(FAILED|ERROR)" gpurun_out/r2g_tests.log | cut -c1-300
