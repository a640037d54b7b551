#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -m gpu -x -q -k "conv or mini" > gpurun_out/t35.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t35.log
tail -4 gpurun_out/t35.log
timeout 300 python tests/conv0_bench.py > gpurun_out/conv0_planned.log 2>&1; tail -3 gpurun_out/conv0_planned.log
timeout 300 python tests/conv0_bench.py --hw 224 --k 3 --filters 32 > gpurun_out/conv0_planned_mb.log 2>&1; tail -2 gpurun_out/conv0_planned_mb.log
