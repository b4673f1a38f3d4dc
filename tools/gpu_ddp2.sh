#!/bin/bash
# 2-GPU check of the round-end bench arms: DDP training (fused optimizer + kernel loss) and inference replicas
cd /root/repo; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 400 $TR bench.py --mode train --gpus 2 --steps 10 --warmup 3 2>&1 | grep -E '^\{|Error|error' | tail -3 | cut -c1-400
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | grep -E '^\{|Error|error' | tail -2 | cut -c1-300
B200_DDP_OPTS=bucket_view timeout 400 $TR bench.py --mode train --gpus 2 --steps 10 --warmup 3 2>&1 | grep -E '^\{|Error|error' | tail -2 | cut -c1-300
