#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -k "pointwise or golden or mini" > gpurun_out/t18.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t18.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=|Error" gpurun_out/t18.log | head -20
timeout 600 python tests/pw_sweep.py 64 10=0,3 2>&1 | tee gpurun_out/pw_sweep_r1w.log
