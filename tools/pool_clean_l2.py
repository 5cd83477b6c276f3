#!/usr/bin/env python
"""Pool kernel rate with a CLEAN vs a DIRTY L2.  In the block's steady state the pool kernel runs right after the
un-pool kernel, whose last ~100 MB of stores are still dirty in the 126 MB L2: the pool's reads evict them, so the
kernel's wall time covers its own 168 MB of reads PLUS that write-back.  Here the same launch is timed (a) after a large
read-only pass that leaves L2 clean and (b) after a large write that leaves it dirty."""
import os, sys, statistics
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import mingraph_unet_b200 as mg
dev = torch.device("cuda:0")
B, C, H, W = 16, 20, 512, 512
fm = torch.randn(B, C, H, W, device=dev).bfloat16()
big_r = torch.randn(300 * 1000 * 1000 // 2, device=dev).bfloat16()
big_w = torch.empty(300 * 1000 * 1000 // 2, device=dev, dtype=torch.bfloat16)
def run(prep):
    ts = []
    for i in range(25):
        prep(); torch.cuda._sleep(200_000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); mg.ops.pool_patches(fm, 16, 16); b.record(); b.synchronize()
        if i >= 5: ts.append(a.elapsed_time(b))
    return statistics.median(ts) * 1e3
nbytes = fm.numel() * 2 + B * 1024 * C * 2
for name, prep in (("clean L2 (after a 300 MB read-only pass)", lambda: torch.sum(big_r)),
                   ("dirty L2 (after a 300 MB fill)", lambda: big_w.fill_(1.0))):
    us = run(prep)
    print(f"pool 168 MB, {name}: {us:6.1f} us = {nbytes / us / 1e3:6.0f} GB/s algorithmic", flush=True)
