#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-variants --no-cpu-baseline > gpurun_out/r04a_bench.json 2> gpurun_out/r04a_bench.err; tail -3 gpurun_out/r04a_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r04a_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
print(json.dumps(d['roofline'], indent=1)[:1800])
PY
