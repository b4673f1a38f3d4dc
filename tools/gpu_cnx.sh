#!/bin/bash
# ConvNeXt training-path check: kernel + model gradient parity, then train / infer throughput with per-op split
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_convnext.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/cnx_tests.log
timeout 300 python tools/bench_model.py convnext_base --batch 128 --steps 10 --warmup 3 --train --detail > gpurun_out/cnx_train.log 2>&1
timeout 300 python tools/bench_model.py convnext_base --batch 256 --steps 10 --warmup 3 --detail > gpurun_out/cnx_infer.log 2>&1
cat gpurun_out/cnx_tests.log gpurun_out/cnx_train.log gpurun_out/cnx_infer.log
