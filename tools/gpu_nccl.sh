#!/bin/bash
# DDP training at N GPUs with different DDP options
N=${1:-2}
cd /root/repo; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for v in "" "bucket_view" "bucket_view,cap=100" "bucket_view,cap=100,static" "cap=200"; do
  echo "== B200_DDP_OPTS=$v $(B200_DDP_OPTS=$v timeout 400 $TR bench.py --mode train --gpus $N --steps 10 --warmup 3 2>&1 | grep -E '^\{' | tail -1 | cut -c80-200)"
done
echo "== single $(timeout 400 python bench.py --mode train --steps 10 --warmup 3 2>&1 | grep -E '^\{' | tail -1 | cut -c80-200)"
