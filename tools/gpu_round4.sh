#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -k 10 $t "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$? $(tail -n 3 gpurun_out/$name.log | cut -c1-900 | tr '\n' ' ')" | tee -a gpurun_out/summary.txt
}
run kernels_bwd 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -k "bwd or dual or residual_gradient"
run kernels 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --no-header -p no:cacheprovider
run train 900 python -m pytest tests/test_gpu_train.py -m gpu -q --no-header -p no:cacheprovider
run models 900 python -m pytest tests/test_gpu_models.py -m gpu -q --no-header -p no:cacheprovider
run bench_train 600 python bench.py --mode train --steps 10 --warmup 3
cat gpurun_out/summary.txt
