#!/bin/bash
# Model-level parity, smoke, bench line, ncu launch list + one full capture of the dominant kernel.
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -k 10 $t "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$? $(tail -n 2 gpurun_out/$name.log | cut -c1-600 | tr '\n' ' ')" | tee -a gpurun_out/summary.txt
}
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run models 900 python -m pytest tests/test_gpu_models.py -m gpu -q -x --no-header -p no:cacheprovider
run kernels 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --no-header -p no:cacheprovider
run bench 600 python bench.py --steps 20 --warmup 5
run bench_ref 600 python bench.py --impl reference --steps 5 --warmup 1
KREG='regex:gemm_bf16|layernorm_|attn_fwd|patchify|assemble_tokens'
run ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 267 -c 178 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline
run ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 150 -c 5 -o gpurun_out/prof_gemm python bench.py --steps 2 --warmup 3 --no-cpu-baseline
run ncu_attn 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 36 -c 1 -o gpurun_out/prof_attn python bench.py --steps 2 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary.txt
