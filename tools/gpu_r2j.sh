#!/bin/bash
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention_bwd" 2>&1 | tail -3 > gpurun_out/r2j_trace.log
timeout 200 python tools/attn_bwd_trace.py >> gpurun_out/r2j_trace.log 2>&1
cat gpurun_out/r2j_trace.log
