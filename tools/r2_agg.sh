#!/bin/bash
# round 2: tensor-core aggregation kernel — parity cases each in its own short-lived process, then the sweep points
mkdir -p gpurun_out
for c in "4096 0 40 64 0 f32" "4099 8 8 64 0 bf16" "4097 3 9 64 1 f32" "4096 17 33 32 0 f32" "9000 1 3 48 1 f32" "70000 8 8 64 0 bf16"; do
  echo "== $c: $(timeout 40 python tools/agg_check.py $c 2>&1 | tail -1)"
done
for v in 1 0; do
  MG_GAT_AGG_MMA=$v timeout 120 python tools/sweep.py --dtype bf16 --graph random --no-ref --points 262144:8:64,262144:32:64,65536:8:64 --out gpurun_out/r2_agg_v$v.md > /dev/null 2>&1
  tail -3 gpurun_out/r2_agg_v$v.md
done
