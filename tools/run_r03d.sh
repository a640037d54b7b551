#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r03d_tests.log 2>&1; tail -3 gpurun_out/r03d_tests.log
t0=$(date +%s); timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r03d_bench.json 2> gpurun_out/r03d_bench.err; t1=$(date +%s); echo "bench wall $((t1-t0)) s"; cut -c1-200 gpurun_out/r03d_bench.json
t0=$(date +%s); timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r03d_ref.json 2> gpurun_out/r03d_ref.err; t1=$(date +%s); echo "reference arm wall $((t1-t0)) s"; cut -c1-300 gpurun_out/r03d_ref.json
