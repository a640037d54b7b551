#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "conv_tma or conv_tcgen05" > gpurun_out/r03o_tests.log 2>&1; tail -3 gpurun_out/r03o_tests.log
for nb in 2 3 4; do echo "== nb $nb"; timeout 300 python tests/kernel_bench.py --only conv3x3_wgrad --knobs 26=$nb 2>&1 | tail -1; done
