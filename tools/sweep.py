#!/usr/bin/env python
"""Graph-layer micro-benchmark sweep (BASELINE.json configs[3]).

    python tools/sweep.py [--quick] [--dtype f32|bf16] [--graph random|grid|both] [--no-ref] [--out FILE]

For every (N, k, F, heads) point: one multi-head GAT layer (in = out = F, heads averaged) on
  * a random graph with fixed in-degree k  (tgt = arange(N).repeat_interleave(k), src = randint, seed 0), and
  * the 4-connected grid graph of the same N (locality contrast; k is ignored),
through ``mingraph_unet_b200.ops.gat_forward`` (the C ABI), timed with CUDA events on the launching
stream after warm-up, an L2 flush (write of a 256 MB buffer) before every timed launch.
Reported per point: ms, edges/s, compulsory GB/s = (N*in*b + N*out*b + 4*(E+N+1) + 8*heads*N) / t and its
fraction of the measured HBM peak, the SpMM gather-model GB/s (E*F*b + N*F*b + 4*(E+N+1)) / t, and the
same layer through the reference's op sequence in stock PyTorch on the same GPU ("torch scatter path":
index -> cat -> linear -> leaky_relu -> max -> exp -> scatter_add_ -> div -> scatter_add_ -> elu,
model/gat/graph_attention.py:53-118, one head at a time, :151), capped where its 28*E*F bytes per head
exceed the memory budget.  The torch path is a throughput comparator written here, not product code.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def torch_scatter_layer(x, src, tgt, Ws, As, slope=0.2):
    """The reference's op sequence for a 1-layer, head-averaging GATNetwork (eval)."""
    N = x.shape[0]
    outs = []
    for W, a in zip(Ws, As):
        Wh = x @ W.t()
        e = torch.nn.functional.leaky_relu(torch.cat([Wh[src], Wh[tgt]], 1) @ a.view(-1, 1), slope).squeeze(-1)
        p = torch.exp(e - e.max())
        den = torch.zeros(N, device=x.device).scatter_add_(0, tgt, p)
        alpha = p / (den[tgt] + 1e-10)
        F = Wh.shape[1]
        hp = torch.zeros(N, F, device=x.device).scatter_add_(0, tgt.unsqueeze(-1).repeat(1, F), alpha.unsqueeze(-1) * Wh[src])
        outs.append(torch.nn.functional.elu(hp))
    return torch.stack(outs, 0).mean(0)


def time_fn(fn, flush, iters, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts), min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--graph", default="both", choices=["random", "grid", "both"])
    ap.add_argument("--heads", type=int, default=4)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default="")
    ap.add_argument("--points", default="", help="explicit N:k:F list, comma separated")
    args = ap.parse_args()

    import mingraph_unet_b200 as mg
    from mingraph_unet_b200 import ops
    dev = torch.device("cuda:0")
    peak, peak_src = measured_peak()
    dt = torch.float32 if args.dtype == "f32" else torch.bfloat16
    b = 4 if dt == torch.float32 else 2
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)        # 256 MB > 126 MB L2

    if args.points:
        pts = [tuple(int(v) for v in p.split(":")) for p in args.points.split(",")]
    elif args.quick:
        pts = [(1024, 8, 64), (16384, 8, 64), (65536, 16, 128), (262144, 8, 64), (262144, 32, 64), (262144, 8, 512)]
    else:
        pts = [(N, k, F) for N in (1024, 4096, 16384, 65536, 262144) for k in (8, 16, 32) for F in (64, 128, 256, 512)]
    kinds = ["random", "grid"] if args.graph == "both" else [args.graph]
    rows = []
    seen_grid = set()
    for N, k, F in pts:
        for kind in kinds:
            if kind == "grid":
                if (N, F) in seen_grid:
                    continue
                seen_grid.add((N, F))
                side = int(round(N ** 0.5))
                if side * side != N:
                    continue
                g = mg.Graph.grid(side, side, dev, 1)
                rowptr, col, E = g.rowptr_in, g.col_in, g.E
                src = tgt = None
            else:
                gen = torch.Generator().manual_seed(0)
                tgt = torch.arange(N).repeat_interleave(k)
                src = torch.randint(0, N, (N * k,), generator=gen)
                ei = torch.stack([src, tgt]).to(dev)
                rowptr, col, _ = ops.csr_from_coo(ei, N, by_target=True)
                E = N * k
                src, tgt = ei[0], ei[1]
            gen = torch.Generator().manual_seed(1)
            x32 = torch.randn(N, F, generator=gen).to(dev)
            x = x32.to(dt)
            H = args.heads
            bound = 1.414 * (6.0 / (F + F)) ** 0.5
            W = ((torch.rand(H, F, F, generator=gen) * 2 - 1) * bound).to(dev)
            bound_a = 1.414 * (6.0 / (2 * F + 1)) ** 0.5
            a = ((torch.rand(H, 2 * F, generator=gen) * 2 - 1) * bound_a).to(dev)
            fn = lambda: ops.gat_forward(x, rowptr, col, W, a, concat=False, slope=0.2, out_dtype=dt)  # noqa: E731
            try:
                ms, ms_min = time_fn(fn, flush, args.iters)
            except Exception as ex:  # noqa: BLE001
                rows.append(dict(N=N, k=k, F=F, graph=kind, error=str(ex)[:120]))
                print(json.dumps(rows[-1]), flush=True)
                continue
            comp = N * F * b + N * F * b + 4 * (E + N + 1) + 8 * H * N
            gath = E * F * b + N * F * b + 4 * (E + N + 1)
            row = dict(N=N, k=(k if kind == "random" else 4), F=F, heads=H, graph=kind, dtype=args.dtype, E=E, ms=ms, ms_min=ms_min,
                       edges_per_s=E / (ms * 1e-3), compulsory_mb=comp / 1e6, compulsory_gbs=comp / (ms * 1e-3) / 1e9,
                       frac_of_peak=comp / (ms * 1e-3) / 1e9 / peak, gather_model_gbs=gath / (ms * 1e-3) / 1e9,
                       flops=2.0 * N * F * H * F, tflops=2.0 * N * F * H * F / (ms * 1e-3) / 1e12)
            if not args.no_ref and kind == "random":
                need = 28.0 * E * F + 8.0 * N * F * 4
                if need < 60e9:
                    Ws, As = [W[h] for h in range(H)], [a[h] for h in range(H)]
                    try:
                        with torch.no_grad():
                            y_ref = torch_scatter_layer(x32, src, tgt, Ws, As)
                            y = fn().float()
                            row["max_abs_vs_torch"] = float((y - y_ref).abs().max())
                            del y_ref, y
                            rms, _ = time_fn(lambda: torch_scatter_layer(x32, src, tgt, Ws, As), flush, max(3, args.iters // 3), warm=1)
                        row["torch_scatter_ms"] = rms
                        row["speedup_vs_torch_scatter"] = rms / ms
                    except torch.OutOfMemoryError:
                        row["torch_scatter_ms"] = None
                        torch.cuda.empty_cache()
                else:
                    row["torch_scatter_ms"] = None      # would materialise > 60 GB of (E,F) intermediates
            rows.append(row)
            print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            f.write(f"GAT layer sweep, dtype {args.dtype}, heads {args.heads}, HBM peak {peak:.0f} GB/s ({peak_src}); "
                    f"L2 flushed before every timed launch; median of {args.iters}\n\n")
            f.write("| graph | N | k | F | E | ms | Medge/s | compulsory MB | compulsory GB/s | frac of peak | gather-model GB/s | "
                    "TFLOP/s (transform) | torch scatter ms | speed-up | max-abs vs torch |\n")
            f.write("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
            for r in rows:
                if "error" in r:
                    f.write(f"| {r['graph']} | {r['N']} | {r['k']} | {r['F']} | error: {r['error']} |\n")
                    continue
                ts = r.get("torch_scatter_ms")
                f.write(f"| {r['graph']} | {r['N']} | {r['k']} | {r['F']} | {r['E']} | {r['ms']:.3f} | {r['edges_per_s'] / 1e6:.0f} | "
                        f"{r['compulsory_mb']:.1f} | {r['compulsory_gbs']:.0f} | {r['frac_of_peak']:.3f} | {r['gather_model_gbs']:.0f} | "
                        f"{r['tflops']:.1f} | {'' if ts is None else f'{ts:.2f}'} | "
                        f"{'' if ts is None else '%.0fx' % r['speedup_vs_torch_scatter']} | "
                        f"{'' if 'max_abs_vs_torch' not in r else '%.1e' % r['max_abs_vs_torch']} |\n")


if __name__ == "__main__":
    main()
