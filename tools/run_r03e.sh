#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/fold_once.py && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dw3x3_rows_fwd|dw_stats_finalize|bn_fold|tc_gemm_kernel|dw_chan_bwd" --launch-skip 14 -o gpurun_out/r03e_fold_unit -f python tools/fold_once.py > gpurun_out/r03e_ncu.log 2>&1; tail -3 gpurun_out/r03e_ncu.log; ls -la gpurun_out/r03e_fold_unit.ncu-rep
