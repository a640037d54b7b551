#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02y_tests.log 2>&1; tail -4 gpurun_out/r02y_tests.log
timeout 500 python bench.py --per-kernel gpurun_out/r02y_r18_perkernel.json > gpurun_out/r02y_r18_bench.json 2> gpurun_out/r02y_r18.err; tail -3 gpurun_out/r02y_r18.err; cut -c1-200 gpurun_out/r02y_r18_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02y_ref_arm.json 2> gpurun_out/r02y_ref_arm.err; tail -2 gpurun_out/r02y_ref_arm.err; cut -c1-400 gpurun_out/r02y_ref_arm.json
timeout 400 python tests/kernel_bench.py --out gpurun_out/r02y_cfg2_kernel_bench.json > gpurun_out/r02y_cfg2_kernel_bench.log 2>&1; tail -25 gpurun_out/r02y_cfg2_kernel_bench.log
timeout 300 python bench.py --workload mnist --no-cpu-baseline --per-kernel gpurun_out/r02y_mnist_perkernel.json > gpurun_out/r02y_mnist_bench.json 2> gpurun_out/r02y_mnist.err; cut -c1-200 gpurun_out/r02y_mnist_bench.json
timeout 400 python bench.py --workload mobilenet --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r02y_mobilenet_bench.json 2> gpurun_out/r02y_mobilenet.err; tail -2 gpurun_out/r02y_mobilenet.err; cut -c1-200 gpurun_out/r02y_mobilenet_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02y_launches.csv python bench.py --no-cpu-baseline --no-e2e --no-variants --no-graph --steps 3 --warmup 3 > gpurun_out/r02y_ncu.log 2>&1; tail -c 200 gpurun_out/r02y_ncu.log
