#!/bin/bash
# round 2 ncu evidence: launch lists of one training step / one inference step, full captures of one block's kernels
# (the reports are converted to CSV on the box: gpurun_out/ may carry 64 MiB back)
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
NCU="ncu --clock-control none --profile-from-start off"
timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_train.csv python tools/ncu_train_step.py > gpurun_out/ncu_train_list.log 2>&1
echo "train list rc=$?"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_infer.csv python tools/ncu_train_step.py --infer > gpurun_out/ncu_infer_list.log 2>&1
echo "infer list rc=$?"
timeout 900 $NCU --set full -o /tmp/prof_train_block -f python tools/ncu_train_step.py --depth 1 > gpurun_out/ncu_train_block.log 2>&1
echo "block full rc=$?"
ncu -i /tmp/prof_train_block.ncu-rep --page raw --csv > gpurun_out/prof_train_block_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -8; du -sh gpurun_out
