#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { # name, batchargs, env...
  name=$1; shift; bargs=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $bargs --warmup 3 --no-cpu-baseline > gpurun_out/bench2p_$name.json 2> gpurun_out/bench2p_$name.err
  python - gpurun_out/bench2p_$name.json $name <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[2], round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms/step', 'e2e', d.get('e2e') and round(d['e2e']['value']), d.get('clocks'))
except Exception as e: print(sys.argv[2], 'ERR', e)
PY
}
run s20_tail "--steps 20"
run s20_none "--steps 20" BENCH_NO_CLOCKS=1
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r1at.json 2> gpurun_out/bench_r1at.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1at.json')); print('1gpu', round(d['value']), d['ms_per_step'], d['e2e']['value'], d['clocks'])"
