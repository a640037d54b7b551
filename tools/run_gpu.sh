#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "residual_join or batchnorm" > gpurun_out/t29.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t29.log
tail -4 gpurun_out/t29.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r1an.json 2> gpurun_out/bench_r1an.err; cut -c1-300 gpurun_out/bench_r1an.json; tail -3 gpurun_out/bench_r1an.err
