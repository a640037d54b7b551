#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { n=$1; shift; tag=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$n bench.py --gpus $n --no-cpu-baseline "$@" > gpurun_out/r03x_${tag}_${n}gpu.json 2> gpurun_out/r03x_${tag}_${n}gpu.err; python -c "
import json,sys
d=json.load(open('gpurun_out/r03x_${tag}_${n}gpu.json')); print('${tag}', d['n_gpus'], round(d['value']), round(d['ms_per_step'],4), d['e2e'] and round(d['e2e']['value']))" || tail -3 gpurun_out/r03x_${tag}_${n}gpu.err; }
timeout 300 python bench.py --no-cpu-baseline --no-variants > gpurun_out/r03x_cfg3_1gpu.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r03x_cfg3_1gpu.json')); print('cfg3', 1, round(d['value']), round(d['ms_per_step'],4))"
run 8 cfg3 --no-variants
run 4 cfg3 --no-variants
