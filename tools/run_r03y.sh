#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python tests/kernel_bench.py --batch 64 --only pw_fwd,pw_dgrad,pw_wgrad 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread --clock-control none -s 6 -c 6 --csv --log-file gpurun_out/r03y_pw_wgrad_b64.csv python tests/kernel_bench.py --batch 64 --only pw_wgrad --iters 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r03y_pw_wgrad_b64.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[h]
for r in rows[h+1:]:
    d=dict(zip(hdr,r))
    print(d['Kernel Name'][:60], d['Metric Name'], d['Metric Value'], d['Metric Unit'])
PY
