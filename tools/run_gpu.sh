#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -k "depthwise" > gpurun_out/t16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t16.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=|Error" gpurun_out/t16.log | head -20
timeout 600 python tests/dw_sweep.py 64 2>&1 | tee gpurun_out/dw_sweep_r1v.log
