#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r03g_tests.log 2>&1; tail -3 gpurun_out/r03g_tests.log
timeout 600 python bench.py --per-kernel gpurun_out/r03g_r18_perkernel.json > gpurun_out/r03g_bench.json 2> gpurun_out/r03g_bench.err; tail -2 gpurun_out/r03g_bench.err; cut -c1-200 gpurun_out/r03g_bench.json
timeout 600 python tests/pw_sweep.py 64 > gpurun_out/r03g_pw_sweep.log 2>&1; head -12 gpurun_out/r03g_pw_sweep.log
