#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "pointwise or conv" > gpurun_out/r02w_tests.log 2>&1; tail -3 gpurun_out/r02w_tests.log
timeout 600 python tests/pw_sweep.py 64 23=0,1 wgrad > gpurun_out/r02w_pw_wgrad_fused_reduce.log 2>&1; cat gpurun_out/r02w_pw_wgrad_fused_reduce.log
