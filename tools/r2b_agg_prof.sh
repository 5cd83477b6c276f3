#!/bin/bash
# round 2 (second session): baseline sweep point, pruned edge-max at k = 8, and one --set full capture with source
mkdir -p gpurun_out
P="--dtype bf16 --graph random --no-ref --points 262144:8:64,262144:16:64,262144:32:64,65536:8:64"
timeout 200 python tools/sweep.py $P --out gpurun_out/r2b_base.md > /dev/null 2>&1; tail -5 gpurun_out/r2b_base.md
MG_GAT_PRUNE_MIN_DEG=8 timeout 200 python tools/sweep.py $P --out gpurun_out/r2b_prune8.md > /dev/null 2>&1; tail -5 gpurun_out/r2b_prune8.md
CMD="python tools/sweep.py --dtype bf16 --graph random --no-ref --iters 2 --points 262144:8:64"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gat_agg_mma|tc_edge_max|tc_scores|tc_u_kernel" -s 8 -c 4 -o gpurun_out/prof_r2b_agg -f $CMD > gpurun_out/r2b_agg_ncu.log 2>&1
tail -2 gpurun_out/r2b_agg_ncu.log
