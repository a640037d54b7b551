#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "folded_into_pointwise or pointwise" > gpurun_out/r02n_tests_a.log 2>&1; tail -5 gpurun_out/r02n_tests_a.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e --per-kernel gpurun_out/r02n_r18_perkernel.json > gpurun_out/r02n_r18_bench.json 2> gpurun_out/r02n_r18.err; tail -3 gpurun_out/r02n_r18.err; cut -c1-250 gpurun_out/r02n_r18_bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02n_r18_perkernel.json'))
for k,v in d['by_call_shape'].items():
    if 'affine' in k: print(k, v['calls'], round(1e3*v['ms']/v['calls'],1),'us')
PY
