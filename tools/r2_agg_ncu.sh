#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/sweep.py --dtype bf16 --graph random --no-ref --iters 2 --points 262144:8:64"
timeout 120 $CMD > gpurun_out/r2_agg_plain.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gat_agg_mma|tc_edge_max|tc_scores|tc_u_kernel" -s 8 -c 4 -o gpurun_out/prof_r2_agg -f $CMD > gpurun_out/r2_agg_ncu.log 2>&1
tail -3 gpurun_out/r2_agg_ncu.log
