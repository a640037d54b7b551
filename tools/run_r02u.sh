#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02u_tests.log 2>&1; tail -4 gpurun_out/r02u_tests.log
timeout 400 python bench.py --per-kernel gpurun_out/r02u_r18_perkernel.json > gpurun_out/r02u_r18_bench.json 2> gpurun_out/r02u_r18.err; tail -3 gpurun_out/r02u_r18.err; cut -c1-250 gpurun_out/r02u_r18_bench.json
timeout 600 python tools/wgrad_traffic.py > gpurun_out/r02u_wgrad_calls.json 2> gpurun_out/r02u_wgrad_calls.err && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel|splitk_reduce|repack_planes" -o gpurun_out/r02u_wgrad_full -f python tools/wgrad_traffic.py > gpurun_out/r02u_ncu.log 2>&1; tail -3 gpurun_out/r02u_ncu.log; ls -la gpurun_out/r02u_wgrad_full.ncu-rep
