#!/bin/bash
# one `ncu --set full` capture per hot-path kernel (third launch of each op at its production shape)
cd /root/repo; mkdir -p gpurun_out
cap() { # name regex op [skip]
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" -s ${4:-2} -c 1 -o gpurun_out/prof_$1 -f python tools/ncu_ops.py $3 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$? $(tail -n 1 gpurun_out/ncu_$1.log | cut -c1-120)"
}
cap ln_fwd layernorm_fwd_kernel ln_fwd
cap ln_bwd layernorm_bwd_bf16 ln_bwd
cap attn_bwd attn_bwd2_kernel attn_bwd
cap attn_long attn_fwd_long attn_long
cap dwconv_wgrad dwconv_wgrad_tma dwconv_wgrad
cap rope rope_qk_kernel rope
cap patchify patchify_kernel patchify
cap assemble assemble_tokens assemble
cap colsum colsum_kernel colsum
cap gemm_dgelu gemm_bf16 gemm_dgelu
cap gemm_wgrad gemm_bf16 gemm_wgrad
