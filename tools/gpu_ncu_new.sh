#!/bin/bash
# `ncu --set full` of the kernels added late in round 1 (third launch of each op at a production shape)
cd /root/repo; mkdir -p gpurun_out
cap() { # name regex op [skip]
  timeout 200 ncu --set full --clock-control none --import-source on -k "regex:$2" -s ${4:-2} -c 1 -o gpurun_out/prof_$1 -f python tools/ncu_ops.py $3 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$? $(tail -n 1 gpurun_out/ncu_$1.log | cut -c1-120)"
}
cap qk_norm qk_norm_fwd_kernel qk_norm 1
cap qk_norm_bwd qk_norm_bwd_kernel qk_norm_bwd
cap glu glu_fwd_kernel glu
cap grn_dot grn_dot_kernel grn 1
cap grn_apply grn_apply_kernel grn 1
cap grn_bwd_apply grn_bwd_apply_kernel grn_bwd
cap patchify_u8 patchify_u8_kernel patchify_u8
cap ema ema_update_kernel ema
