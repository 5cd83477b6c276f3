#!/bin/bash
mkdir -p gpurun_out
for c in 5 6 7; do
  MG_BLOCK_CLUSTER=$c timeout 60 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-extras > gpurun_out/r2_cl$c.log 2> gpurun_out/r2_cl$c.err
  python - gpurun_out/r2_cl$c.log $c <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    o = l["roofline"]["other_kernels"]
    print("cluster %s: step %.1f us | latency %.1f us | block %.1f us | e2e_nf %d" % (sys.argv[2], l["ms_per_step"]*1e3, l["step_latency_ms"]*1e3, o["block_forward_kernel"]["ms"]*1e3, l["e2e_node_features"]["value"]))
except Exception as e:
    print("cluster", sys.argv[2], "FAILED", e, open(sys.argv[1].replace(".log", ".err")).read()[-300:])
PY
done
