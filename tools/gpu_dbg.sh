#!/bin/bash
mkdir -p gpurun_out; cd $GRAFT_REPO_ROOT
timeout 120 python tools/debug_patch.py > gpurun_out/dbg.log 2>&1; cat gpurun_out/dbg.log | cut -c1-600
