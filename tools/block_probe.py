#!/usr/bin/env python
"""Probe: fused block kernel time vs batch size (clusters of 8 CTAs): does it step when the clusters stop being co-resident?"""
import os, sys, statistics
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import mingraph_unet_b200 as mg
dev = torch.device("cuda:0")
blk = mg.GraphBlock(node_feature_dim=20, num_segments=2).to(dev).eval()
prep = blk._prepared()
for hp in (32, 64):
    for B in (1, 4, 8, 12, 14, 15, 16, 17, 18, 20, 24, 32):
        x = torch.randn(B, hp * hp, 20, device=dev).to(torch.bfloat16)
        ts = []
        for i in range(30):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(100_000)
            a.record()
            mg.ops.block_forward(x, hp, hp, prep, 64, 4, 2, 4, 2)
            b.record()
            b.synchronize()
            if i >= 5:
                ts.append(a.elapsed_time(b))
        print(f"grid {hp}x{hp} B={B:3d}: {1e3 * statistics.median(ts):7.1f} us", flush=True)
