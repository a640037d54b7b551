#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02v_tests.log 2>&1; tail -4 gpurun_out/r02v_tests.log
timeout 600 python tests/pw_sweep.py 64 23=0,1 wgrad > gpurun_out/r02v_pw_wgrad_fused_reduce.log 2>&1; cat gpurun_out/r02v_pw_wgrad_fused_reduce.log
timeout 400 python bench.py --no-cpu-baseline --per-kernel gpurun_out/r02v_r18_perkernel.json > gpurun_out/r02v_r18_bench.json 2> gpurun_out/r02v_r18.err; tail -3 gpurun_out/r02v_r18.err; cut -c1-250 gpurun_out/r02v_r18_bench.json
