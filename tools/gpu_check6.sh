#!/bin/bash
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
( timeout 300 python tools/wgrad_bench.py ) > gpurun_out/r2m_wgrad.log 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r2m_tests.log 2>&1
( timeout 300 python tools/bench_model.py convnext_base --batch 256 --detail ) > gpurun_out/r2m_convnext_detail.log 2>&1
( timeout 300 python tools/ln_fold_ab.py ) > gpurun_out/r2m_fold_ab.log 2>&1
cat gpurun_out/r2m_wgrad.log; tail -6 gpurun_out/r2m_tests.log; tail -30 gpurun_out/r2m_convnext_detail.log; head -5 gpurun_out/r2m_fold_ab.log
