#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload cfg5 --steps 2 --warmup 1"
timeout 200 $CMD > gpurun_out/r2_train_plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train.csv $CMD > gpurun_out/r2_train_ncu.log 2>&1
tail -c 300 gpurun_out/r2_train_plain.log; wc -l gpurun_out/r2_launches_train.csv
# large-map pool rate (size effect of the 168 MB read)
timeout 100 python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
import mingraph_unet_b200 as mg
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for B, H in ((16, 512), (64, 512), (64, 1024)):
    x = torch.randn(B, 20, H, H, device="cuda").bfloat16()
    ts = []
    for i in range(15):
        flush.zero_(); torch.cuda._sleep(200000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); mg.ops.pool_patches(x, 16); b.record(); torch.cuda.synchronize()
        if i >= 5: ts.append(a.elapsed_time(b))
    ts.sort(); t = ts[len(ts) // 2]
    print("pool B=%d %dx%d: %.1f MB in %.1f us = %.0f GB/s" % (B, H, H, x.numel() * 2 / 1e6, t * 1e3, x.numel() * 2 / t / 1e6))
PY
