#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "conv or pointwise" > gpurun_out/t24.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t24.log
tail -3 gpurun_out/t24.log
timeout 300 python tests/pw_sweep.py 64 11=0,1 wgrad > gpurun_out/pw_sweep_shortA.log 2>&1; tail -12 gpurun_out/pw_sweep_shortA.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r1ag.json 2> gpurun_out/bench_r1ag.err; cut -c1-400 gpurun_out/bench_r1ag.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_rows" -c 4 -f -o gpurun_out/prof_conv0_r1ag python tests/conv0_bench.py --once > gpurun_out/ncu_conv0_r1ag.log 2>&1
tail -2 gpurun_out/ncu_conv0_r1ag.log
