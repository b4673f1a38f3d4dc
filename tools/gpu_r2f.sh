#!/bin/bash
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514"
timeout 240 $TR tools/ddp_parity.py --out gpurun_out/r2f_ddp_parity.json > gpurun_out/r2f_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/r2f_parity.log
timeout 240 $TR tools/ddp_detail.py > gpurun_out/r2f_ddp_detail.log 2>&1
echo "detail rc=$?" >> gpurun_out/r2f_ddp_detail.log
grep -v "Warning\|warn\|^$\|\*\*\*\|OMP_NUM" gpurun_out/r2f_parity.log | tail -12 | cut -c1-1500
grep -v "Warning\|warn\|^$\|\*\*\*\|OMP_NUM" gpurun_out/r2f_ddp_detail.log | tail -14 | cut -c1-400
