#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r03z_tests.log 2>&1; tail -3 gpurun_out/r03z_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r03z_bench.json 2> gpurun_out/r03z_bench.err; tail -2 gpurun_out/r03z_bench.err; cut -c1-200 gpurun_out/r03z_bench.json
timeout 300 python tests/kernel_bench.py > gpurun_out/r03z_cfg2_kernels.txt 2>&1; tail -13 gpurun_out/r03z_cfg2_kernels.txt
timeout 300 python tests/kernel_bench.py --batch 64 > gpurun_out/r03z_cfg2_kernels_batch64.txt 2>&1; tail -13 gpurun_out/r03z_cfg2_kernels_batch64.txt
