#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --no-variants --per-kernel gpurun_out/r03s_r18_perkernel.json > gpurun_out/r03s_bench.json 2> gpurun_out/r03s_bench.err; tail -2 gpurun_out/r03s_bench.err; cut -c1-200 gpurun_out/r03s_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r03s_launches.csv python bench.py --no-cpu-baseline --no-e2e --no-variants --no-graph --steps 3 --warmup 3 > gpurun_out/r03s_ncu.log 2>&1; tail -c 150 gpurun_out/r03s_ncu.log
