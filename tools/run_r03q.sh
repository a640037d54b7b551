#!/bin/bash
cd $GRAFT_REPO_ROOT
for d in 15 8207 10255; do
  echo "dbg=$d $(timeout 60 python tests/kernel_bench.py --only conv3x3_wgrad --knobs 27=$d 2>&1 | tail -1)"
done
for d in 271 8463; do
  echo "=== dbg=$d"
  timeout 60 python tests/kernel_bench.py --only conv3x3_wgrad --iters 1 --knobs 27=$d > /tmp/o.txt 2>&1
  for r in producer mma shifter; do grep "cw2 cta 73 $r" /tmp/o.txt | tail -1; done
done
