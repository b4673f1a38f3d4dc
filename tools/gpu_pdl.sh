#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -4
for m in 0 1 0 1; do
  echo "== B200_PDL=$m $(B200_PDL=$m timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c80-190)"
done
B200_PDL=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-graph 2>&1 | tail -1 | cut -c80-190
B200_PDL=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-graph 2>&1 | tail -1 | cut -c80-190
timeout 300 python tools/gap_probe.py
