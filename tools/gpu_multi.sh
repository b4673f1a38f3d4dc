#!/bin/bash
# N-GPU runs (launched exactly as the driver launches them): inference replicas and DDP training.
N=${1:-2}
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -k 10 $t "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$? $(grep -E '^\{' gpurun_out/$name.log | tail -n 1 | cut -c1-420) $(tail -n 2 gpurun_out/$name.log | grep -v '^{' | cut -c1-300 | tr '\n' ' ')" | tee -a gpurun_out/summary.txt
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run infer_n1 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline
run infer_n$N 600 $TR bench.py --gpus $N --steps 20 --warmup 5
run train_n1 400 python bench.py --mode train --gpus 1 --steps 10 --warmup 3
run train_n$N 600 $TR bench.py --mode train --gpus $N --steps 10 --warmup 3
run ref_n$N 600 $TR bench.py --impl reference --gpus $N --steps 3 --warmup 1
NCCL_DEBUG=INFO timeout 300 $TR bench.py --mode train --gpus $N --steps 2 --warmup 3 2>&1 | grep -E "NVLS|Using network|via P2P|Channel 00.*via" | head -8 > gpurun_out/nccl_info.txt
cat gpurun_out/summary.txt
