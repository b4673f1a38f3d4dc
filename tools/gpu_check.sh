#!/bin/bash
# full GPU regression + headline benches after a kernel change
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/check_tests.log; cat gpurun_out/check_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/check_bench.log 2>&1; tail -1 gpurun_out/check_bench.log | cut -c1-400
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/check_train.log 2>&1; tail -1 gpurun_out/check_train.log | cut -c1-300
timeout 300 python tools/microbench.py > gpurun_out/check_micro.log 2>&1; cat gpurun_out/check_micro.log | cut -c1-200
timeout 300 python tools/bench_model.py convnext_base --batch 256 --steps 10 --warmup 3 2>&1 | tail -1
