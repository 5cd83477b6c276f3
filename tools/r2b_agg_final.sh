#!/bin/bash
# round 2 (second session): parity of the reworked aggregation kernel, sweep points, one --set full capture
mkdir -p gpurun_out
for c in "4099 0 40 64 0 f32" "70000 8 8 64 0 bf16" "5000 0 9 48 1 bf16" "4097 1 3 16 0 bf16"; do
  echo "check $c: $(timeout 60 python tools/agg_check.py $c 2>&1 | tail -1)"
done
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu 2>&1 | tail -2
P="--dtype bf16 --graph random --no-ref --points 262144:8:64,262144:16:64,262144:32:64,65536:8:64,65536:32:64,16384:16:64"
timeout 200 python tools/sweep.py $P --out gpurun_out/r2b_sweep_f64_bf16.md > /dev/null 2>&1; tail -7 gpurun_out/r2b_sweep_f64_bf16.md | cut -d'|' -f3,4,7,11,12
CMD="python tools/sweep.py --dtype bf16 --graph random --no-ref --iters 2 --points 262144:8:64"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gat_agg_mma|tc_edge_max|tc_scores|tc_u_kernel" -s 8 -c 4 -o gpurun_out/prof_r2b_agg_final -f $CMD > gpurun_out/r2b_agg_final_ncu.log 2>&1
tail -1 gpurun_out/r2b_agg_final_ncu.log
