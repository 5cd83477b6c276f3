#!/bin/bash
# round 2: tensor-core aggregation kernel — parity cases each in its own short-lived process, then the sweep points
mkdir -p gpurun_out
for c in "4096 0 40 64 0 f32" "4099 8 8 64 0 bf16" "70000 8 8 64 0 bf16"; do
  echo "== $c: $(timeout 40 python tools/agg_check.py $c 2>&1 | tail -1)"
done
timeout 200 python -m pytest tests/test_gpu_tc.py -x -q -m gpu 2>&1 | tail -2
timeout 120 python tools/sweep.py --dtype bf16 --graph random --no-ref --points 262144:8:64,262144:16:64,262144:32:64,65536:8:64 --out gpurun_out/r2_agg_v1.md > /dev/null 2>&1
tail -4 gpurun_out/r2_agg_v1.md
