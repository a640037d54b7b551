#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tests/pw_sweep.py 1 > gpurun_out/r03c_pw_sweep_batch1.log 2>&1; cat gpurun_out/r03c_pw_sweep_batch1.log | head -20
timeout 600 python tests/bn_stats_sweep.py 4 | head -5
