#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench4_r1az.json 2> gpurun_out/bench4_r1az.err
python -c "
import json; d=json.load(open('gpurun_out/bench4_r1az.json')); print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), d['clocks']['mode'][-60:])"
