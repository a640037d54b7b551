#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "packed_operands" > gpurun_out/r03k_tests.log 2>&1; tail -12 gpurun_out/r03k_tests.log
