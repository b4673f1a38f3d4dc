#!/bin/bash
# round 2, 2-GPU call: DDP graph capture, DDP parity, scaling of the train step
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/ddp_parity.py --out gpurun_out/r2b_ddp_parity.json > gpurun_out/r2b_parity.log 2>&1
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 600 $TR bench.py --gpus 2 --mode train --steps 10 --warmup 3 > gpurun_out/r2b_train_n2.log 2>&1
grep -E "NVLS|comm 0x|Init COMPLETE|nranks" gpurun_out/r2b_train_n2.log | head -20 > gpurun_out/r2b_nccl.txt
grep -c "AllReduce" gpurun_out/r2b_train_n2.log >> gpurun_out/r2b_nccl.txt
timeout 600 $TR bench.py --gpus 2 --mode train --steps 10 --warmup 3 --no-graph > gpurun_out/r2b_train_n2_eager.log 2>&1
timeout 600 python bench.py --gpus 1 --mode train --steps 10 --warmup 3 --no-cpu-baseline --no-script-records > gpurun_out/r2b_train_n1.log 2>&1
tail -3 gpurun_out/r2b_parity.log; grep '^{' gpurun_out/r2b_train_n2.log | cut -c1-900; tail -2 gpurun_out/r2b_train_n2_eager.log | cut -c1-900; tail -1 gpurun_out/r2b_train_n1.log | cut -c1-600; cat gpurun_out/r2b_nccl.txt | cut -c1-200
