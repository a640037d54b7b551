#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "folded_into_pointwise or unit_fused" > gpurun_out/r02o_tests_a.log 2>&1; tail -12 gpurun_out/r02o_tests_a.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02o_tests.log 2>&1; tail -25 gpurun_out/r02o_tests.log
timeout 300 python bench.py --no-cpu-baseline --per-kernel gpurun_out/r02o_r18_perkernel.json > gpurun_out/r02o_r18_bench.json 2> gpurun_out/r02o_r18.err; tail -3 gpurun_out/r02o_r18.err; cut -c1-250 gpurun_out/r02o_r18_bench.json
