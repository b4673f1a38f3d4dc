#!/bin/bash
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout 240 $TR tools/ddp_detail.py > gpurun_out/r2d_ddp_detail.log 2>&1
grep -v "Warning\|warn" gpurun_out/r2d_ddp_detail.log | tail -40
