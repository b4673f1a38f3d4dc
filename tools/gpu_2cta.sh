#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.log
GEMM_TRACE_DUMP=1 B200_GEMM_2CTA=1 timeout 200 python tools/gemm_trace.py > gpurun_out/trace_2cta.log 2>&1; grep -E "events|quantiles" gpurun_out/trace_2cta.log
GEMM_TRACE_DUMP=1 timeout 200 python tools/gemm_trace.py > gpurun_out/trace_1cta.log 2>&1; grep -E "events|quantiles" gpurun_out/trace_1cta.log
