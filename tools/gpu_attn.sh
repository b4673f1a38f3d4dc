#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
for m in 0 4 6; do
echo "== poly $m"; B200_ATTN_LONG_POLY=$m timeout 600 python tools/bench_model.py naflex:vit_large_patch14_clip_336 --batch 64 --steps 10 --warmup 3 --detail 2>&1 | grep -E "attention|img_per_s" | cut -c1-200
done
echo "== v1"; B200_ATTN_FWD_V1=1 timeout 600 python tools/bench_model.py naflex:vit_large_patch14_clip_336 --batch 64 --steps 10 --warmup 3 --detail 2>&1 | grep -E "attention|img_per_s" | cut -c1-200
