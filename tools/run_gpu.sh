#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t33.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t33.log
tail -12 gpurun_out/t33.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r1ar.json 2> gpurun_out/bench_r1ar.err; cut -c1-300 gpurun_out/bench_r1ar.json; tail -3 gpurun_out/bench_r1ar.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1ar.json')); print(d['final_loss'], d['e2e'], d['gpu_launches'])"
