#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r1av.json 2> gpurun_out/bench_r1av.err; tail -3 gpurun_out/bench_r1av.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1av.json')); print(round(d['value']), d['ms_per_step'], d['e2e'], d['clocks'])"
