#!/bin/bash
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_plugin.py -m gpu -x -q -k "mask" > gpurun_out/dbg.log 2>&1; tail -30 gpurun_out/dbg.log | cut -c1-400
