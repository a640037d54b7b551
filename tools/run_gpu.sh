#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t37.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t37.log
tail -8 gpurun_out/t37.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r1ba.json 2> gpurun_out/bench_r1ba.err; tail -3 gpurun_out/bench_r1ba.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1ba.json')); print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['final_loss'], d['gpu_launches'])"
