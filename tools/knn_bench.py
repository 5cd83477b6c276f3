#!/usr/bin/env python
"""kNN graph build timing (csrc/knn.cu): per-image graphs of the named configs and a few larger single graphs.
FP32 pipe work = 3 flops (sub, mul, add; no FMA by design: oracle parity) per (pair, feature)."""
import json
import os
import statistics
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import mingraph_unet_b200 as mg  # noqa: E402

cases = [(16, 1024, 64, 8), (16, 1024, 64, 32), (64, 4096, 64, 8), (8, 4096, 64, 16), (1, 16384, 64, 8), (1, 16384, 128, 16),
         (1, 65536, 64, 8)]
if len(sys.argv) > 1 and sys.argv[1] == "--quick":
    cases = cases[:3]
print("| graphs | nodes/graph | D | k | ms | Gpair-dims/s | FP32 TFLOP/s (3 flop) | Medge/s |\n|---:|---:|---:|---:|---:|---:|---:|---:|")
for B, npg, D, k in cases:
    x = torch.randn(B * npg, D, device="cuda")
    for _ in range(2):
        mg.ops.knn_graph(x, k, nodes_per_graph=npg if B > 1 else 0)
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        mg.ops.knn_graph(x, k, nodes_per_graph=npg if B > 1 else 0)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ms = statistics.median(ts)
    pd = B * npg * npg * D
    print(f"| {B} | {npg} | {D} | {k} | {ms:.3f} | {pd / ms / 1e6:.1f} | {3 * pd / ms / 1e9:.1f} | {B * npg * k / ms / 1e3:.1f} |", flush=True)
