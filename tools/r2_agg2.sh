#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python tools/sweep.py --dtype bf16 --graph random --no-ref --points 262144:8:64,262144:16:64,262144:32:64,65536:8:64,65536:32:64,16384:16:64 --out gpurun_out/r2_sweep_f64_bf16.md > /dev/null 2>&1
tail -7 gpurun_out/r2_sweep_f64_bf16.md
CMD="python tools/sweep.py --dtype bf16 --graph random --no-ref --iters 2 --points 262144:8:64"
timeout 120 $CMD > gpurun_out/r2_agg_plain.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gat_agg_mma|tc_edge_max|tc_scores|tc_u_kernel" -s 8 -c 4 -o gpurun_out/prof_r2_agg2 -f $CMD > gpurun_out/r2_agg_ncu.log 2>&1
tail -2 gpurun_out/r2_agg_ncu.log
