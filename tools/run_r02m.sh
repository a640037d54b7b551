#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02m_tests.log 2>&1; tail -15 gpurun_out/r02m_tests.log
timeout 300 python bench.py --no-cpu-baseline --per-kernel gpurun_out/r02m_r18_perkernel.json > gpurun_out/r02m_r18_bench.json 2> gpurun_out/r02m_r18.err; tail -3 gpurun_out/r02m_r18.err; cut -c1-300 gpurun_out/r02m_r18_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02m_launches.csv python bench.py --no-cpu-baseline --no-e2e --no-graph --steps 3 --warmup 3 > gpurun_out/r02m_ncu.log 2>&1; tail -c 300 gpurun_out/r02m_ncu.log
