#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { # name, ngpu, env...
  name=$1; shift; ng=$1; shift
  if [ $ng = 1 ]; then
    env "$@" CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/benchc_$name.json 2> gpurun_out/benchc_$name.err
  else
    env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/benchc_$name.json 2> gpurun_out/benchc_$name.err
  fi
  python - gpurun_out/benchc_$name.json $name <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[2], round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms/step', d.get('clocks'))
except Exception as e: print(sys.argv[2], 'ERR', e)
PY
}
run g1_inline 1 BENCH_CLOCK_MODE=inline
run g1_thread 1 BENCH_CLOCK_MODE=thread
run g1_none 1 BENCH_NO_CLOCKS=1
run g2_inline 2 BENCH_CLOCK_MODE=inline
run g2_thread 2 BENCH_CLOCK_MODE=thread
run g2_none 2 BENCH_NO_CLOCKS=1
