#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_network.py -x -q -m gpu -k "autograph or unchanged" > gpurun_out/r03n_tests.log 2>&1; tail -12 gpurun_out/r03n_tests.log
