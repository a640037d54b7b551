#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tests/bn_stats_sweep.py 1,2,4,8,16 > gpurun_out/r03b_bn_stats_sweep.log 2>&1; cat gpurun_out/r03b_bn_stats_sweep.log
