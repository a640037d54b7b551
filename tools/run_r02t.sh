#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tests/pw_sweep.py 64 22=256,128,64,0 wgrad > gpurun_out/r02t_pw_wgrad_bn_sweep.log 2>&1; cat gpurun_out/r02t_pw_wgrad_bn_sweep.log
