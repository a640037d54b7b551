#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tests/bn_shapes_probe.py 16 > gpurun_out/bnshapes.log 2>&1; cat gpurun_out/bnshapes.log
timeout 600 python tests/bn_shapes_probe.py 64 > gpurun_out/bnshapes64.log 2>&1; cat gpurun_out/bnshapes64.log
