#!/bin/bash
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "layernorm or rmsnorm" 2>&1 | tail -3 > gpurun_out/r2k.log
for m in 1 2; do echo "mode $m" >> gpurun_out/r2k.log; B200_LN_BWD_RING=$m timeout 100 python tools/ln_bench.py >> gpurun_out/r2k.log 2>&1; done
cat gpurun_out/r2k.log
