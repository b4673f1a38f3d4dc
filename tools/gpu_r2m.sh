#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
timeout 300 $TR tools/ddp_parity.py --modes peer --out gpurun_out/r2m_parity_n$N.json > gpurun_out/r2m_parity_n$N.log 2>&1
echo "parity rc=$?" >> gpurun_out/r2m_parity_n$N.log
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --mode train > gpurun_out/r2m_bench_n$N.log 2>&1
echo "bench rc=$?" >> gpurun_out/r2m_bench_n$N.log
grep -E '^\{|rc=' gpurun_out/r2m_parity_n$N.log | cut -c1-900
grep -E '^\{|rc=|Error|error' gpurun_out/r2m_bench_n$N.log | cut -c1-400
