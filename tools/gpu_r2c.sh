#!/bin/bash
# 2-GPU: all-reduce in isolation; DDP bucket / NCCL CTA variants of the graphed train step
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 120 $TR tools/nccl_bench.py 2>&1 | grep '^{' > gpurun_out/r2c_nccl_bench.log
run() { echo "== $1" >> gpurun_out/r2c_variants.log; env $1 timeout 150 $TR bench.py --gpus 2 --mode train --steps 10 --warmup 3 2>&1 | grep '^{' | cut -c100-260 >> gpurun_out/r2c_variants.log; }
run "B200_DDP_OPTS="
run "B200_DDP_OPTS=cap=1024"
run "B200_DDP_OPTS=cap=60"
run "NCCL_MAX_CTAS=4"
run "NCCL_MAX_CTAS=8"
run "NCCL_NVLS_ENABLE=0"
run "NCCL_ALGO=Ring"
cat gpurun_out/r2c_nccl_bench.log gpurun_out/r2c_variants.log
