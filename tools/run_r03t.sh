#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw_chan_bwd -s 3 -c 1 -o gpurun_out/r03t_dw_chan_bwd -f python tests/kernel_bench.py --only dw_bwd --iters 2 > gpurun_out/r03t_ncu1.log 2>&1; tail -2 gpurun_out/r03t_ncu1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_s1_wgrad2 -s 3 -c 1 -o gpurun_out/r03t_conv_wgrad2 -f python tests/kernel_bench.py --only conv3x3_wgrad --iters 2 > gpurun_out/r03t_ncu2.log 2>&1; tail -2 gpurun_out/r03t_ncu2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_s1_kernel -s 3 -c 1 -o gpurun_out/r03t_conv_fwd -f python tests/kernel_bench.py --only conv3x3_fwd --iters 2 > gpurun_out/r03t_ncu3.log 2>&1; tail -2 gpurun_out/r03t_ncu3.log
ls -la gpurun_out/*.ncu-rep | tail -3
