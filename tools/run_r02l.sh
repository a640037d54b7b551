#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "folded_into_pointwise" > gpurun_out/r02l_tests_a.log 2>&1; tail -15 gpurun_out/r02l_tests_a.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02l_tests.log 2>&1; tail -15 gpurun_out/r02l_tests.log
timeout 300 python bench.py --no-cpu-baseline --per-kernel gpurun_out/r02l_r18_perkernel.json > gpurun_out/r02l_r18_bench.json 2> gpurun_out/r02l_r18.err; tail -3 gpurun_out/r02l_r18.err; cut -c1-300 gpurun_out/r02l_r18_bench.json
