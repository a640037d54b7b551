#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "strided_pointwise" > gpurun_out/t31.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t31.log
tail -4 gpurun_out/t31.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-graph --per-kernel gpurun_out/perkernel_r1ap.json > gpurun_out/bench_r1ap_eager.json 2> gpurun_out/bench_r1ap.err; tail -2 gpurun_out/bench_r1ap.err
