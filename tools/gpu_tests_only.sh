#!/bin/bash
# usage: bash tools/gpu_tests_only.sh [pytest -k expression]   -- the GPU suite (or a selection of it) on the box
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
if [ -n "$1" ]; then
  ( time timeout 600 python -m pytest tests -m gpu -q -k "$1" 2>&1 | tail -25 ) > gpurun_out/gpu_tests.log 2>&1
else
  ( time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/gpu_tests.log 2>&1
fi
tail -12 gpurun_out/gpu_tests.log
