#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py -x -q -m gpu -k "conv_tma or conv_tcgen05 or mnist" > gpurun_out/r03q_tests.log 2>&1; tail -3 gpurun_out/r03q_tests.log
for k in "26=0" "26=2"; do
  echo "knobs=$k $(timeout 60 python tests/kernel_bench.py --only conv3x3_wgrad --knobs $k 2>&1 | tail -1)"
done
