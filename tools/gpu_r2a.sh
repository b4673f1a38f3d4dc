#!/bin/bash
# round 2, first GPU call: plugin on real timm, graphed train step, the new bench line
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_smi.txt
nproc > gpurun_out/r2a_cpu.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r2a_cpu.txt 2>&1; python -c "import os; print(len(os.sched_getaffinity(0)))" >> gpurun_out/r2a_cpu.txt
timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_loss.py tests/test_gpu_plugin.py -q -x 2>&1 | tail -40 > gpurun_out/r2a_tests.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2a_bench.log 2>&1
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/r2a_ref.log 2>&1
tail -5 gpurun_out/r2a_tests.log; tail -c 3000 gpurun_out/r2a_bench.log; tail -c 1500 gpurun_out/r2a_ref.log
