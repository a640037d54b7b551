#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r03j_tests.log 2>&1; tail -12 gpurun_out/r03j_tests.log
for v in 0 1; do DK_PW_PACK=$v timeout 500 python bench.py --no-cpu-baseline --no-variants --no-e2e > gpurun_out/r03j_bench_pack$v.json 2> gpurun_out/r03j_pack$v.err; tail -2 gpurun_out/r03j_pack$v.err; python -c "
import json; d=json.load(open('gpurun_out/r03j_bench_pack$v.json')); print('pack $v', round(d['value']), round(d['ms_per_step'],4), d['final_loss'], d['gpu_launches']//20, d['roofline']['kernel'], round(d['roofline']['frac'],3))"; done
