#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bn_fused_bwd|bn_bwd_" -s 60 -c 12 -f -o gpurun_out/prof_bnbwd_r1af python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/ncu_bnbwd_r1af.log 2>&1
tail -2 gpurun_out/ncu_bnbwd_r1af.log
