#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py -x -q -m gpu -k "conv_tma or conv_tcgen05 or mnist" > gpurun_out/r03p_tests.log 2>&1; tail -3 gpurun_out/r03p_tests.log
timeout 300 python tests/kernel_bench.py --only conv3x3_wgrad,conv3x3_fwd 2>&1 | tail -2
