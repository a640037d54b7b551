#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02p_tests.log 2>&1; tail -25 gpurun_out/r02p_tests.log
timeout 300 python bench.py --no-cpu-baseline --per-kernel gpurun_out/r02p_r18_perkernel.json > gpurun_out/r02p_r18_bench.json 2> gpurun_out/r02p_r18.err; tail -3 gpurun_out/r02p_r18.err; cut -c1-250 gpurun_out/r02p_r18_bench.json
