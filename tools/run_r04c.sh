#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py tests/test_gpu_golden.py -x -q -m gpu -k "bn or batch or join or residual or network or resnet or golden" 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --no-variants --no-e2e --per-kernel gpurun_out/r04c_perkernel.json 2>/dev/null | cut -c1-230
python - <<'PY'
import json
d=json.load(open('gpurun_out/r04c_perkernel.json'))
for k,v in d['by_call_shape'].items():
    if 'join' in k: print(k, v)
PY
