#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -k 10 $t "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$? $(tail -n 2 gpurun_out/$name.log | cut -c1-800 | tr '\n' ' ')" | tee -a gpurun_out/summary.txt
}
run kernels 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --no-header -p no:cacheprovider
run models 900 python -m pytest tests/test_gpu_models.py -m gpu -q --no-header -p no:cacheprovider
run microbench 600 python tools/microbench.py
run bench 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
cat gpurun_out/summary.txt
