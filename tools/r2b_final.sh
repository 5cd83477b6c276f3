#!/bin/bash
# round 2, second session, one GPU: full parity suite, the bench line, ncu launch lists, the configs[3] sweeps
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/r2b_bench_n1.log 2> gpurun_out/r2b_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2b_bench_n1.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r2b_bench_n1.log").read().strip().splitlines()[-1])
    o = l["roofline"]["other_kernels"]
    print("step %.1f us | %d img/s | latency %s | pool %.1f us | block %.1f | unpool %.1f | e2e %d | e2e_nf %d | cpu %s" % (
        l["ms_per_step"] * 1e3, l["value"], l["step_latency_ms"], o["pool_patches_tma_kernel"]["ms"] * 1e3,
        o["block_forward_kernel"]["ms"] * 1e3, l["roofline"]["kernel_ms"] * 1e3, l["e2e"]["value"], l["e2e_node_features"]["value"],
        l.get("cpu_baseline")))
except Exception as e:
    print("FAILED", e)
PY
CMD="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras"
timeout 200 $CMD > gpurun_out/r2b_plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches.csv $CMD > gpurun_out/r2b_ncu_l.log 2>&1
wc -l gpurun_out/r2b_launches.csv
timeout 500 python tools/sweep.py --dtype bf16 --out gpurun_out/r2b_sweep_bf16.md > gpurun_out/r2b_sweep_bf16.log 2>&1; tail -1 gpurun_out/r2b_sweep_bf16.log | cut -c1-200
timeout 500 python tools/sweep.py --dtype f32 --out gpurun_out/r2b_sweep_f32.md > gpurun_out/r2b_sweep_f32.log 2>&1; tail -1 gpurun_out/r2b_sweep_f32.log | cut -c1-200
SW="python tools/sweep.py --dtype bf16 --graph random --no-ref --iters 3 --points 262144:8:64,262144:32:64,262144:8:128,262144:8:512"
timeout 200 $SW > gpurun_out/r2b_sweep_plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_sweep.csv $SW > gpurun_out/r2b_ncu_sw.log 2>&1
wc -l gpurun_out/r2b_launches_sweep.csv
