#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_data.py tests/test_gpu_convnext.py tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -5
cap() { timeout 200 ncu --set full --clock-control none --import-source on -k "regex:$2" -s ${4:-2} -c 1 -o gpurun_out/prof_$1 -f python tools/ncu_ops.py $3 > gpurun_out/ncu_$1.log 2>&1; echo "$1 rc=$?"; }
cap patchify_u8 patchify_u8_kernel patchify_u8
cap grn_apply grn_apply_kernel grn 1
cap grn_bwd_apply grn_bwd_apply_kernel grn_bwd
