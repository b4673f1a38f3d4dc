#!/bin/bash
# Full evidence run: GPU test suite, smoke, bench (infer + train), reference arm, ncu launch list and full captures.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -k 10 $t "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$? $(tail -n 2 gpurun_out/$name.log | cut -c1-2500 | tr '\n' ' ')" | tee -a gpurun_out/summary.txt
}
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run tests 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider
run bench 600 python bench.py --steps 20 --warmup 5
run bench_train 600 python bench.py --mode train --steps 10 --warmup 3
run bench_ref 600 python bench.py --impl reference --steps 5 --warmup 1
run microbench 600 python tools/microbench.py
run convnext 300 python tools/bench_model.py convnext_base --detail
run convnext_train 300 python tools/bench_model.py convnext_base --batch 128 --steps 10 --warmup 3 --train --detail
run train_detail 300 python tools/train_detail.py
run gemm_trace 300 python tools/gemm_trace.py
run attn_trace 300 python tools/attn_trace.py
KREG='regex:gemm_bf16|layernorm_|attn_|patchify|assemble_tokens'
run ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 356 -c 178 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph
run ncu_gemm 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 150 -c 5 -o gpurun_out/prof_gemm python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph
run ncu_attn 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd2 -s 36 -c 1 -o gpurun_out/prof_attn python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph
run ncu_dwconv 600 ncu --set full --clock-control none --import-source on -k regex:dwconv_tma -s 12 -c 1 -o gpurun_out/prof_dwconv python tools/bench_model.py convnext_base --batch 256 --steps 1 --warmup 1
cat gpurun_out/summary.txt
