#!/bin/bash
# final-tree ncu evidence of round 2: launch list of one inference forward (LayerNorm folded into the GEMMs), full captures of
# every kernel of a depth-2 inference forward and of the kernels added late in the round (reports -> CSV on the box)
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
NCU="ncu --clock-control none --profile-from-start off"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_infer_b.csv python tools/ncu_train_step.py --infer > gpurun_out/ncu_infer_list_b.log 2>&1
echo "infer list rc=$?"
timeout 900 $NCU --set full -o /tmp/prof_infer2 -f python tools/ncu_train_step.py --infer --depth 2 > gpurun_out/ncu_infer2.log 2>&1
echo "infer depth-2 full rc=$?"
ncu -i /tmp/prof_infer2.ncu-rep --page raw --csv > gpurun_out/prof_infer2_raw.csv 2>/dev/null
timeout 600 $NCU --set full -o /tmp/prof_new -f python tools/ncu_new_kernels.py > gpurun_out/ncu_new.log 2>&1
echo "new kernels full rc=$?"
ncu -i /tmp/prof_new.ncu-rep --page raw --csv > gpurun_out/prof_new_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -8; du -sh gpurun_out
