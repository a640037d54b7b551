#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02z_tests.log 2>&1; tail -4 gpurun_out/r02z_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 500 python bench.py --no-cpu-baseline --no-variants --per-kernel gpurun_out/r02z_r18_perkernel.json > gpurun_out/r02z_r18_bench.json 2> gpurun_out/r02z_r18.err; tail -3 gpurun_out/r02z_r18.err; cut -c1-200 gpurun_out/r02z_r18_bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02z_r18_perkernel.json'))
for k,v in d['by_call_shape'].items():
    if 'dwconv_bwd' in k or 'join' in k: print(k, v['calls'], round(1e3*v['ms']/v['calls'],1),'us')
PY
