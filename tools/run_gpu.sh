#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "batchnorm" > gpurun_out/t25.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t25.log
tail -8 gpurun_out/t25.log
timeout 300 python tests/bn_sweep.py 64 1,2 > gpurun_out/bn_sweep_r1ah.log 2>&1; grep -E "HW=14|HW=7" gpurun_out/bn_sweep_r1ah.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r1ah.json 2> gpurun_out/bench_r1ah.err; cut -c1-300 gpurun_out/bench_r1ah.json; tail -3 gpurun_out/bench_r1ah.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1ah.json')); print(d['clocks'], d['e2e'])
PY
