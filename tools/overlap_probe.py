#!/usr/bin/env python
"""Probe: does splitting the cfg-2 batch over concurrently replayed CUDA graphs (one per stream) raise throughput?
The block kernel is latency-bound (69 us, one CTA per SM), pool / un-pool are HBM-bound; independent half batches on two
streams let them overlap.  Prints images/s for 1, 2 and 4 concurrent shards of the same 16 images."""
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import mingraph_unet_b200 as mg  # noqa: E402

H = W = 512
B, C, D = 16, 20, 64
dev = torch.device("cuda:0")
blk = mg.GraphBlock(node_feature_dim=C, num_segments=2).to(dev).eval()
fm = torch.randn(B, C, H, W, device=dev).to(torch.bfloat16)
fusion = torch.zeros(B, 32 + D, H, W, dtype=torch.bfloat16, device=dev)
for shards in (1, 2, 4):
    per = B // shards
    streams = [torch.cuda.Stream() for _ in range(shards)]
    runners = []
    for i in range(shards):
        with torch.cuda.stream(streams[i]):
            runners.append(mg.CapturedGraphBlock(blk, fm[i * per:(i + 1) * per].contiguous(), image_size=(H, W),
                                                 out=fusion[i * per:(i + 1) * per, 32:]))
    torch.cuda.synchronize()

    def step():
        for i in range(shards):
            with torch.cuda.stream(streams[i]):
                runners[i]()

    for _ in range(20):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    t0 = time.perf_counter()
    K = 300
    for _ in range(K):
        step()
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / K
    print(f"shards {shards}: {ms * 1e3:.1f} us per 16 images, {B / ms * 1e3:.0f} images/s (wall {1e3 * (time.perf_counter() - t0) / K:.3f} ms/step)")
