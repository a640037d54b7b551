#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_network.py -x -q -m gpu -k p2p > gpurun_out/r02r_p2p_test.log 2>&1; tail -6 gpurun_out/r02r_p2p_test.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r02r_bench_2gpu.json 2> gpurun_out/r02r_bench_2gpu.err; tail -3 gpurun_out/r02r_bench_2gpu.err; cut -c1-250 gpurun_out/r02r_bench_2gpu.json
timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r02r_bench_1gpu.json 2> gpurun_out/r02r_bench_1gpu.err; cut -c1-250 gpurun_out/r02r_bench_1gpu.json
