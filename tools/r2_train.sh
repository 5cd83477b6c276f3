#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_backward.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python bench.py --workload cfg5 --steps 100 --warmup 5 > gpurun_out/r2_train.log 2> gpurun_out/r2_train.err; tail -c 400 gpurun_out/r2_train.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r2_train.log").read().strip().splitlines()[-1])
print("train step %.3f ms | %d img/s | launches/step %s | bwd kernel %.1f us (%.2f) | e2e %d" % (l["ms_per_step"], l["value"], l["launch_mode"][-12:], l["roofline"]["kernel_ms"]*1e3, l["roofline"]["frac"], l["e2e"]["value"]))
PY
