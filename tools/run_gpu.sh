#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/t22.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t22.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=|Error|assert" gpurun_out/t22.log | head -20
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --per-kernel gpurun_out/perkernel_r1ad.json > gpurun_out/bench_r1ad.json 2> gpurun_out/bench_r1ad.err
cut -c1-300 gpurun_out/bench_r1ad.json; tail -3 gpurun_out/bench_r1ad.err
