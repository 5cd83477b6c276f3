#!/bin/bash
# round 2, one GPU: parity of everything, then the bench (no ncu)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/r2_bench_n1.log 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r2_bench_n1.log").read().strip().splitlines()[-1])
    o = l["roofline"]["other_kernels"]
    print("step %.1f us | %d img/s | latency %s | pool %.1f us | block %.1f | unpool %.1f | e2e %d | e2e_nf %d | cpu %s" % (
        l["ms_per_step"] * 1e3, l["value"], l["step_latency_ms"], o["pool_patches_tma_kernel"]["ms"] * 1e3,
        o["block_forward_kernel"]["ms"] * 1e3, l["roofline"]["kernel_ms"] * 1e3, l["e2e"]["value"], l["e2e_node_features"]["value"],
        l.get("cpu_baseline")))
    print(json.dumps(l.get("extra"), indent=1)[:3000])
except Exception as e:
    print("FAILED", e)
PY
