#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r03r_tests.log 2>&1; tail -3 gpurun_out/r03r_tests.log
timeout 300 python tests/kernel_bench.py > gpurun_out/r03r_kernels.txt 2>&1; cat gpurun_out/r03r_kernels.txt | tail -20
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-variants > gpurun_out/r03r_bench.json 2> gpurun_out/r03r_bench.err; cat gpurun_out/r03r_bench.json | cut -c1-600
