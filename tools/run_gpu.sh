#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -m gpu -x -q -k "pointwise or mini or full_size" > gpurun_out/t34.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t34.log
tail -4 gpurun_out/t34.log
timeout 300 python tests/pw_sweep.py 64 15=1,2,4,0 wgrad > gpurun_out/pw_sweep_wide.log 2>&1; grep "s=1" gpurun_out/pw_sweep_wide.log
