#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r03a_tests.log 2>&1; tail -4 gpurun_out/r03a_tests.log
for v in 0 1; do DK_ASYNC_WGRAD=$v timeout 500 python bench.py --no-cpu-baseline --no-variants --no-e2e > gpurun_out/r03a_r18_bench_async$v.json 2> gpurun_out/r03a_r18_async$v.err; tail -2 gpurun_out/r03a_r18_async$v.err; python -c "
import json; d=json.load(open('gpurun_out/r03a_r18_bench_async$v.json')); print('async $v', round(d['value']), round(d['ms_per_step'],4), d['final_loss'])"; done
