#!/usr/bin/env python
"""Error and time of the GAT layer through the 3xTF32 tensor-pipe transform (csrc/gat_tc_gemm.cu) against the FP32-pipe
kernels, both against the CPU oracle.  Runs each path in a fresh process (the switch is read once per process):
    python tests/check_tc_gemm.py            # parent: prints a table
"""
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
SHAPES = [(8192, 4, 64, 64), (8192, 4, 128, 128), (8192, 1, 128, 128), (4096, 4, 256, 256), (4096, 1, 256, 256),
          (4096, 4, 512, 512), (4096, 1, 512, 512), (65536, 4, 128, 128)]


def child():
    import torch
    sys.path.insert(0, ROOT)
    from oracle import restate as O
    import mingraph_unet_b200 as mg
    for N, heads, fin, fout in SHAPES:
        gen = torch.Generator().manual_seed(N + heads + fin)
        k = 8
        tgt = torch.arange(N).repeat_interleave(k)
        src = torch.randint(0, N, (N * k,), generator=gen)
        ei = torch.stack([src, tgt])
        x = torch.randn(N, fin, generator=gen)
        Ws, As = O.init_gat_params(fin, fout, heads, gen)
        rowptr, col, _ = mg.ops.csr_from_coo(ei.cuda(), N, by_target=True)
        xg, Wg, Ag = x.cuda(), Ws.cuda(), As.cuda()
        y = mg.ops.gat_forward(xg, rowptr, col, Wg, Ag, concat=False)
        err = None
        if N <= 8192:
            ref = O.gat_layer(x, ei, Ws, As, 0.2, concat=False)
            err = float((y.cpu() - ref).abs().max())
            mag = float(ref.abs().max())
        else:
            mag = float(y.abs().max())
        ts = []
        for _ in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            mg.ops.gat_forward(xg, rowptr, col, Wg, Ag, concat=False)
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        print(f"{N} {heads} {fin} {fout} {mag:.3f} {err if err is not None else float('nan'):.3e} {min(ts):.4f}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    res = {}
    for tag, env in (("3xTF32", "1"), ("FP32 pipe", "0")):
        out = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, MG_GAT_TC_GEMM=env), capture_output=True, text=True)
        if out.returncode:
            print(out.stderr[-2000:])
        res[tag] = [ln.split() for ln in out.stdout.strip().splitlines()]
    print("| N | heads | in | F | max abs(ref) | err 3xTF32 | err FP32 pipe | ms 3xTF32 | ms FP32 pipe | speed-up |")
    print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for a, b in zip(res["3xTF32"], res["FP32 pipe"]):
        print(f"| {a[0]} | {a[1]} | {a[2]} | {a[3]} | {a[4]} | {a[5]} | {b[5]} | {a[6]} | {b[6]} | {float(b[6]) / float(a[6]):.1f}x |")
