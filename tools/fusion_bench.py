#!/usr/bin/env python
"""Rate of the per-region fusion gather (csrc/fusion.cu, mg_region_map_gather) at the block's sizes: per-pixel region map ->
dense (B, D, H, W) channel slice of a fused buffer.  HBM-write bound: algorithmic bytes = D*H*W*b + H*W*sizeof(label) +
R*D*4 per image.  Label patterns: 'runs' (superpixel-like 8 x 8 cells: the one-read-per-16-bytes path) and 'random'
(every pixel its own label: one float4 table read per pixel and 4 channels).  Compared with torch's own
index + permute + slice-copy of the same result."""
import json
import os
import statistics
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import mingraph_unet_b200 as mg  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timeit(fn, flush):
    ts = []
    for i in range(13):
        flush.fill_(0.0)                       # 256 MB write: L2 holds none of the inputs / outputs
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    dev = torch.device("cuda:0")
    P = peak()
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    print(f"HBM peak {P:.0f} GB/s\n| B | HxW | R | D | out | map | labels | ms | algorithmic GB/s | frac | torch ms |\n|---:|---|---:|---:|---|---|---|---:|---:|---:|---:|")
    for (B, H, W, R, D) in [(16, 512, 512, 2, 64), (16, 512, 512, 300, 64), (8, 1024, 1024, 300, 64)]:
        gen = torch.Generator().manual_seed(R)
        table = torch.randn(R, D, generator=gen).to(dev)
        for pattern in ("runs", "random"):
            if pattern == "runs":
                m64 = torch.randint(0, R, (B, H // 8, W // 8), generator=gen).repeat_interleave(8, 1).repeat_interleave(8, 2)
            else:
                m64 = torch.randint(0, R, (B, H, W), generator=gen)
            for odt, mdt in ((torch.bfloat16, torch.int32), (torch.bfloat16, torch.int64), (torch.float32, torch.int64)):
                m = m64.to(mdt).to(dev)
                fused = torch.empty(B, 32 + D, H, W, dtype=odt, device=dev)
                dst = fused[:, 32:]
                ms = timeit(lambda: mg.ops.region_map_gather(table, m, out=dst), flush)
                nbytes = B * (D * H * W * fused.element_size() + H * W * m.element_size()) + R * D * 4
                tms = timeit(lambda: dst.copy_(table[m.long()].permute(0, 3, 1, 2)), flush)
                gbs = nbytes / ms / 1e6
                print(f"| {B} | {H}x{W} | {R} | {D} | {str(odt)[6:]} | {str(mdt)[6:]} | {pattern} | {ms:.4f} | {gbs:.0f} | "
                      f"{gbs / P:.2f} | {tms:.3f} |", flush=True)


if __name__ == "__main__":
    main()
