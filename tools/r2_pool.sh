#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pool" 2>&1 | tail -2
timeout 150 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-extras > gpurun_out/r2_pool.log 2> gpurun_out/r2_pool.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r2_pool.log").read().strip().splitlines()[-1])
o = l["roofline"]["other_kernels"]
print("step %.1f us | latency %.1f | pool %.1f us (%.2f) | block %.1f | unpool %.1f" % (l["ms_per_step"]*1e3, l["step_latency_ms"]*1e3, o["pool_patches_tma_kernel"]["ms"]*1e3, o["pool_patches_tma_kernel"]["frac"], o["block_forward_kernel"]["ms"]*1e3, l["roofline"]["kernel_ms"]*1e3))
PY
