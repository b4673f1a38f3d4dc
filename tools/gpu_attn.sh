#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q --no-header -p no:cacheprovider -k "attention" 2>&1 | tail -3
for d in 0 1500 2500 3500; do
  echo "== stagger $d"; B200_ATTN_STAGGER=$d timeout 300 python tools/microbench.py 2>&1 | grep attention | cut -c1-140
done
timeout 300 python tools/attn_trace.py 2>&1 | grep -E "launch|softmax|total"
