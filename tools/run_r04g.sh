#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "conv_tma_stride1 or conv_tma_wgrad" > gpurun_out/r04g_memcheck.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r04g_memcheck.log
