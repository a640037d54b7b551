#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_s1_wgrad2 -s 3 -c 1 -o gpurun_out/r04d_conv_wgrad2 -f python tests/kernel_bench.py --only conv3x3_wgrad --iters 2 > gpurun_out/r04d_ncu1.log 2>&1; tail -1 gpurun_out/r04d_ncu1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 3 -c 1 -o gpurun_out/r04d_pw_fwd_7x7 -f python tests/kernel_bench.py --only pw_fwd --batch 64 --chans 512 --hw 7 --iters 2 > gpurun_out/r04d_ncu2.log 2>&1; tail -1 gpurun_out/r04d_ncu2.log
timeout 120 python tests/kernel_bench.py --batch 64 --chans 512 --hw 7 --only pw_fwd,pw_dgrad,pw_wgrad,bn_fwd_train,bn_bwd,dw_fwd,dw_bwd 2>&1 | tail -7
timeout 120 python tests/kernel_bench.py --batch 64 --chans 256 --hw 14 --only pw_fwd,pw_dgrad,pw_wgrad,bn_fwd_train,bn_bwd,dw_fwd,dw_bwd 2>&1 | tail -7
