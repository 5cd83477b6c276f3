#!/bin/bash
# A/B matrix for the variants written after the round-1 GPU budget was spent.  Every command is wrapped in `timeout`;
# results land in gpurun_out/ab_*.log.  Usage (from the repo root):
#   gpurun --timeout 1200 -- 'bash tools/ab_next_round.sh 1'           # one GPU: fusion gather, pool kernel variants (tensor-core sum, dynamic strips)
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/ab_next_round.sh 2'   # two GPUs: exchange variants (inline / captured / bucketed / p2p)
#   gpurun --gpus 8 --timeout 400 -- 'bash tools/ab_next_round.sh 8 inline'   # one variant at a time at 8 GPUs
N=${1:-1}
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    o = l["roofline"]["other_kernels"]
    print(sys.argv[1], "| %.1f us/step | %d images/s | latency %.1f us | pool %.1f us | e2e %d | exchange %s" % (
        l["ms_per_step"] * 1e3, l["value"], l["step_latency_ms"] * 1e3, o["pool_patches_tma_kernel"]["ms"] * 1e3,
        l["e2e"]["value"], l.get("exchange")))
except Exception as e:
    print(sys.argv[1], "FAILED:", e)
PY
}
if [ "$N" = "1" ]; then
  # kernels written after the round-1 GPU budget was spent: parity first, then their rate
  MG_TEST_UNVERIFIED=1 timeout 300 python -m pytest tests/test_gpu_zfusion.py -x -q -m gpu 2>&1 | tail -3
  timeout 200 python tools/fusion_bench.py > gpurun_out/ab_fusion.md 2> gpurun_out/ab_fusion.err; tail -20 gpurun_out/ab_fusion.md
  MG_POOL_DYNAMIC=1 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pool or block" 2>&1 | tail -2
  MG_POOL_MMA=1 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pool or block" 2>&1 | tail -2
  MG_POOL_MMA=1 timeout 150 python bench.py --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/ab_mma.log 2> gpurun_out/ab_mma.err
  summ gpurun_out/ab_mma.log
  for dyn in 0 1; do
    MG_POOL_DYNAMIC=$dyn timeout 150 python bench.py --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/ab_dyn$dyn.log 2> gpurun_out/ab_dyn$dyn.err
    summ gpurun_out/ab_dyn$dyn.log
  done
else
  modes=${2:-"inline captured bucketed p2p"}
  port=29600
  for m in $modes; do
    if [ "$m" != "captured" ] || [ "$N" = "2" ]; then
      port=$((port + 1))
      timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
        tools/check_captured_gather.py --mode $m 2>&1 | grep -E "gather world|rror" | tail -2
    fi
    port=$((port + 1))
    timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --steps 200 --warmup 10 --exchange $m > gpurun_out/ab_n${N}_$m.log 2> gpurun_out/ab_n${N}_$m.err
    echo "rc=$?"; summ gpurun_out/ab_n${N}_$m.log
  done
  # NUMA binding of the ranks (on by default at N>1): the same run without it, for the e2e leg
  port=$((port + 1))
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $N --steps 200 --warmup 10 --no-numa-bind > gpurun_out/ab_n${N}_nonuma.log 2> gpurun_out/ab_n${N}_nonuma.err
  echo "rc=$?"; summ gpurun_out/ab_n${N}_nonuma.log
fi
