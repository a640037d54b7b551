#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py -x -q -m gpu -k "depthwise or dw or network or resnet" > gpurun_out/r03w_tests.log 2>&1; tail -3 gpurun_out/r03w_tests.log
timeout 120 python tests/kernel_bench.py --only dw_bwd,dw_fwd 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline --no-variants --no-e2e 2>/dev/null | cut -c1-230
