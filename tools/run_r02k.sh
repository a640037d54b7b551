#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02k_tests.log 2>&1; tail -8 gpurun_out/r02k_tests.log
timeout 300 python bench.py --workload mnist --no-cpu-baseline --per-kernel gpurun_out/r02k_mnist_perkernel.json > gpurun_out/r02k_mnist_bench.json 2> gpurun_out/r02k_mnist.err; tail -3 gpurun_out/r02k_mnist.err; cut -c1-300 gpurun_out/r02k_mnist_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02k_mnist_launches.csv python bench.py --workload mnist --no-cpu-baseline --no-e2e --no-graph --steps 3 --warmup 3 > gpurun_out/r02k_ncu.log 2>&1; tail -c 300 gpurun_out/r02k_ncu.log
