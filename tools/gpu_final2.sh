#!/bin/bash
# Round-end regression + evidence on the final tree: smoke, the whole GPU suite, the three bench arms, the training-tail
# microbenchmark, the ncu launch list of the bench and a full capture of the optimizer kernel (new this round).
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
run optim_bf16 300 python tools/optim_bench.py --dtype bf16
run optim_fp32 300 python tools/optim_bench.py --dtype fp32
KREG='regex:gemm_bf16|layernorm_|attn_|patchify|assemble_tokens'
run ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 356 -c 178 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph
run ncu_optim 600 ncu --set full --clock-control none --import-source on -k regex:optim_step_kernel -s 6 -c 2 -o gpurun_out/prof_optim python tools/optim_bench.py --dtype bf16 --steps 2
cat gpurun_out/summary.txt
