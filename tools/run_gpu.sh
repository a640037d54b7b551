#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/r01bc_bench.json 2> gpurun_out/r01bc_bench.err; tail -2 gpurun_out/r01bc_bench.err; cut -c1-200 gpurun_out/r01bc_bench.json
