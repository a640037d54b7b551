#!/bin/bash
# round 2, call i: new tests (checkpoint, folded inference, materialised convs), MNIST bench + per-kernel table
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_checkpoint.py tests/test_gpu_network.py -x -q -m gpu -k "checkpoint or folded or mnist" > gpurun_out/r02i_tests_a.log 2>&1; tail -15 gpurun_out/r02i_tests_a.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "conv" > gpurun_out/r02i_tests_b.log 2>&1; tail -8 gpurun_out/r02i_tests_b.log
timeout 300 python bench.py --workload mnist --no-cpu-baseline --per-kernel gpurun_out/r02i_mnist_perkernel.json > gpurun_out/r02i_mnist_bench.json 2> gpurun_out/r02i_mnist.err; tail -3 gpurun_out/r02i_mnist.err; cut -c1-300 gpurun_out/r02i_mnist_bench.json
