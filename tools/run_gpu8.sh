#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 tests/p2p_check.py > gpurun_out/p2p_check_$N.log 2>&1; echo "rc=$?" >> gpurun_out/p2p_check_$N.log; grep "p2p_check\|rc=\|Error\|error" gpurun_out/p2p_check_$N.log | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench${N}_r1au.json 2> gpurun_out/bench${N}_r1au.err
python - gpurun_out/bench${N}_r1au.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(d['n_gpus'], 'gpus', round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms/step', 'e2e', d.get('e2e') and round(d['e2e']['value']), d['config'].get('allreduce')[:30], d.get('clocks'))
except Exception as e: print('ERR', e)
PY
grep -i "error\|trap\|timed out\|unavailable" gpurun_out/bench${N}_r1au.err | head -5
