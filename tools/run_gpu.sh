#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "depthwise" > gpurun_out/t26.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t26.log
tail -8 gpurun_out/t26.log
timeout 300 python tests/dw_sweep.py 64 1,5 > gpurun_out/dw_sweep_r1ai.log 2>&1; grep -E "HW=14 s=1|HW=7 s=1" gpurun_out/dw_sweep_r1ai.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r1ai.json 2> gpurun_out/bench_r1ai.err; cut -c1-300 gpurun_out/bench_r1ai.json; tail -3 gpurun_out/bench_r1ai.err
