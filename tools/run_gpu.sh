#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
DK_HYBRID_WGRAD=1 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -m gpu -x -q -k "pointwise or mini" > gpurun_out/t38.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t38.log
tail -4 gpurun_out/t38.log
timeout 200 python tests/pw_sweep.py 64 16=0,1 wgrad > gpurun_out/pw_sweep_hybrid.log 2>&1; grep "s=1" gpurun_out/pw_sweep_hybrid.log
