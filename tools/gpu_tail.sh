#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_optim.py tests/test_gpu_loss.py -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -12
for i in 1 2; do
  B200_TORCH_OPTIM=1 B200_TORCH_LOSS=1 timeout 300 python bench.py --mode train --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('torch tail', d['value'], d['ms_per_step'], d['gpu_launches'])"
  timeout 300 python bench.py --mode train --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('b200 tail ', d['value'], d['ms_per_step'], d['gpu_launches'])"
done
