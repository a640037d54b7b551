#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
n=$(./tools/tma_probe)
for i in $(seq 0 $((n-1))); do timeout 60 ./tools/tma_probe $i; done > gpurun_out/probe3.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t10.log
timeout 300 python tests/conv0_bench.py --batch 4 > gpurun_out/conv0_b4.log 2>&1
timeout 300 python tests/conv0_bench.py --out gpurun_out/conv0_r1r.json > gpurun_out/conv0_r1r.log 2>&1
timeout 300 python tests/conv0_bench.py --knobs 8=0 > gpurun_out/conv0_r1r_old.log 2>&1
timeout 300 python tests/kernel_bench.py --only pw_dgrad --batch 64 --hw 56 --stride 2 > gpurun_out/kb_pwd_s2.log 2>&1
tail -5 gpurun_out/t10.log; cat gpurun_out/conv0_b4.log gpurun_out/conv0_r1r.log gpurun_out/conv0_r1r_old.log gpurun_out/kb_pwd_s2.log; cat gpurun_out/probe3.log
