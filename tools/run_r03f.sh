#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "pointwise or folded or fused or conv or dense" > gpurun_out/r03f_tests.log 2>&1; tail -3 gpurun_out/r03f_tests.log
timeout 600 python tests/pw_sweep.py 64 25=0,32,64,128 fwd,dgrad > gpurun_out/r03f_pw_bres_sweep.log 2>&1; cat gpurun_out/r03f_pw_bres_sweep.log
