#!/bin/bash
# First GPU bring-up: every kernel group in its own process under `timeout` so a hung kernel cannot eat the box.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { # name, timeout, cmd...
  local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -k 10 $t "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$? $(tail -n 3 gpurun_out/$name.log | tr '\n' ' ')" | tee -a gpurun_out/summary.txt
}
T="python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --no-header -p no:cacheprovider"
run devcheck 120 python -c "import sys; sys.path.insert(0,'pytorch-image-models_b200'); import torch; from b200timm import _lib; l=_lib.load(); print('check', l.b200_device_check(), l.b200_last_error())"
run ln 300 $T -k "layernorm"
run helpers 300 $T -k "patchify or assemble"
run gemm_kk 300 $T -k "gemm_layouts and False-False"
run gemm_kmn 300 $T -k "gemm_layouts and False-True"
run gemm_mnmn 300 $T -k "gemm_layouts and True-True"
run gemm_mnk 300 $T -k "gemm_layouts and True-False"
run gemm_epi 300 $T -k "gemm_epilogues or gemm_fp32 or gemm_vit or gemm_rejects"
run attn 300 $T -k "attention"
run microbench 600 python tools/microbench.py
cat gpurun_out/summary.txt
