#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2 3 4; do
  timeout 600 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/r04h_tests_$i.log 2>&1; echo "run $i: $(tail -1 gpurun_out/r04h_tests_$i.log)"
done
for i in 1 2 3; do timeout 200 python bench.py --no-cpu-baseline --no-variants --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(d['value']), round(d['ms_per_step'],4), d['final_loss'])"; done
