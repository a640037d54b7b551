#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r03u_tests.log 2>&1; tail -3 gpurun_out/r03u_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --per-kernel gpurun_out/r03u_r18_perkernel.json > gpurun_out/r03u_bench.json 2> gpurun_out/r03u_bench.err; tail -2 gpurun_out/r03u_bench.err; cut -c1-200 gpurun_out/r03u_bench.json
timeout 300 python bench.py --workload mnist --no-cpu-baseline > gpurun_out/r03u_mnist_bench.json 2>/dev/null; cut -c1-200 gpurun_out/r03u_mnist_bench.json
timeout 300 python tests/kernel_bench.py > gpurun_out/r03u_cfg2_kernels.txt 2>&1; tail -13 gpurun_out/r03u_cfg2_kernels.txt
timeout 120 tools/micro/mma_rate.bin > gpurun_out/r03u_mma_rate.txt 2>&1
timeout 120 tools/micro/pdl_gap.bin > gpurun_out/r03u_pdl_gap.txt 2>&1
timeout 120 python tools/tf32_peak.py > gpurun_out/r03u_tf32_peak.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r03u_launches.csv python bench.py --no-cpu-baseline --no-e2e --no-variants --no-graph --steps 3 --warmup 3 > gpurun_out/r03u_ncu.log 2>&1; tail -c 150 gpurun_out/r03u_ncu.log
