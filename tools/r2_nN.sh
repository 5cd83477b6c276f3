#!/bin/bash
# round 2, N GPUs: the exchange check at world N (both modes), then the bench with each exchange mode
N=${1:-2}
modes=${2:-"p2p inline"}
mkdir -p gpurun_out
port=29610
for m in $modes; do
  port=$((port + 1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
    tools/check_exchange.py --mode $m 2>&1 | grep -E "exchange world|rror|Traceback" | tail -3
  port=$((port + 1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $N --steps 200 --warmup 10 --exchange $m > gpurun_out/r2_bench_n${N}_$m.log 2> gpurun_out/r2_bench_n${N}_$m.err
  echo "bench $m rc=$?"
  tail -c 800 gpurun_out/r2_bench_n${N}_$m.err
  python - gpurun_out/r2_bench_n${N}_$m.log <<'PY'
import json, sys
try:
    l = json.loads([x for x in open(sys.argv[1]).read().strip().splitlines() if x.startswith("{")][-1])
    x = l.get("extra", {})
    print("N=%d step %.1f us | %d img/s | e2e %d | e2e_nf %d | status %s | cfg3 %s | cfg5 %s" % (
        l["n_gpus"], l["ms_per_step"] * 1e3, l["value"], l["e2e"]["value"], l["e2e_node_features"]["value"], l.get("exchange_status"),
        {k: round(v, 3) for k, v in x.get("cfg3", {}).items() if k in ("images_per_s", "ms_per_step", "efficiency", "ms_per_step_1gpu_same_run")},
        {k: round(v, 3) for k, v in x.get("cfg5", {}).items() if k in ("images_per_s", "ms_per_step", "efficiency", "ms_per_step_1gpu_same_run")}))
except Exception as e:
    print("FAILED", e)
PY
done
