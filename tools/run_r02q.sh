#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for z in 0; do echo "== fold, DK_FOLD_ZEROSUM=$z"; DK_FOLD_ZEROSUM=$z timeout 300 python tests/net_parity.py --net r18 --backend 1 2>&1 | grep -v "^E  \|Traceback\|File " | tail -18; done > gpurun_out/r02q2.log 2>&1
cat gpurun_out/r02q2.log
