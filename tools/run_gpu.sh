#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t36.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t36.log
tail -6 gpurun_out/t36.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r1ax.json 2> gpurun_out/bench_r1ax.err; tail -3 gpurun_out/bench_r1ax.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1ax.json')); print(round(d['value']), d['ms_per_step'], d['e2e'], d['final_loss'])"
