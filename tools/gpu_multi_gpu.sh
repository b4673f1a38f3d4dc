#!/bin/bash
# N-GPU validation of the data-parallel path: parity (DDP + peer push) and the bench line at N ranks
N=${1:-4}
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520"
timeout 300 $TR tools/ddp_parity.py --modes peer --out gpurun_out/r2l_parity_n$N.json > gpurun_out/r2l_parity_n$N.log 2>&1
echo "parity rc=$?" >> gpurun_out/r2l_parity_n$N.log
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2l_bench_n$N.log 2>&1
echo "bench rc=$?" >> gpurun_out/r2l_bench_n$N.log
timeout 200 $TR bench.py --gpus $N --steps 20 --warmup 5 --mode train --dp ddp > gpurun_out/r2l_bench_ddp_n$N.log 2>&1
timeout 100 $TR tools/nccl_bench.py 2>&1 | grep '^{' > gpurun_out/r2l_nccl_n$N.log
grep -E '^\{|rc=' gpurun_out/r2l_parity_n$N.log | cut -c1-700
grep -E '^\{|rc=|Error|error' gpurun_out/r2l_bench_n$N.log | cut -c1-1500
grep -E '^\{' gpurun_out/r2l_bench_ddp_n$N.log | cut -c1-400
cat gpurun_out/r2l_nccl_n$N.log
