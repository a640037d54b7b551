#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"bn_fused_bwd|bn_bwd_|bn_group_bwd" -s 84 -c 28 -o gpurun_out/r03l_bn_bwd_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --no-variants > gpurun_out/r03l_ncu.log 2>&1; tail -2 gpurun_out/r03l_ncu.log; ls -la gpurun_out/r03l_bn_bwd_full.ncu-rep
