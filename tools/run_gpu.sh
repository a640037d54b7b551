#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -m gpu -x -q -k "pointwise or conv or mini or full_size" > gpurun_out/t27.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t27.log
tail -4 gpurun_out/t27.log
timeout 300 python tests/pw_sweep.py 64 14=0,1 fwd,dgrad > gpurun_out/pw_sweep_2cta_b.log 2>&1; grep "s=1" gpurun_out/pw_sweep_2cta_b.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r1ak.json 2> gpurun_out/bench_r1ak.err; cut -c1-300 gpurun_out/bench_r1ak.json; tail -3 gpurun_out/bench_r1ak.err
