#!/bin/bash
# round-end evidence on one B200: GPU tests, bench (ours + reference arm), ncu launch list, full capture of the dominant family
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${1:-r01aw}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-250 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err; cut -c1-200 gpurun_out/${T}_bench_reference_arm.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/${T}_ncu_launches.log 2>&1
wc -l gpurun_out/${T}_ncu_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bn_fused_bwd|bn_group_bwd" -s 66 -c 33 -f -o gpurun_out/${T}_prof_bn_bwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/${T}_ncu_bn_bwd.log 2>&1
tail -1 gpurun_out/${T}_ncu_bn_bwd.log
