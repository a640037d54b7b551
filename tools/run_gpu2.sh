#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 --no-e2e > gpurun_out/bench2_$name.json 2> gpurun_out/bench2_$name.err
  python - gpurun_out/bench2_$name.json $name <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[2], round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms/step', d.get('clocks'))
except Exception as e: print(sys.argv[2], 'ERR', e)
PY
}
run clock BENCH_CLOCK_CALLS=clock
run power BENCH_CLOCK_CALLS=power
run reasons BENCH_CLOCK_CALLS=reasons
run period1 BENCH_CLOCK_PERIOD=1.0
run none BENCH_CLOCK_CALLS=none
