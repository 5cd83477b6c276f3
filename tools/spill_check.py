"""Parity cases for the wide-row tensor-core aggregation (gat_agg_spill_kernel), each printed on its own line:
python tools/spill_check.py"""
import sys, torch
sys.path.insert(0, ".")
from oracle import restate as O
import mingraph_unet_b200 as mg
cases = [  # N, kmin, kmax, in, F, graphs
    # (all with 2 N in F heads >= 1e9, the TMA-GEMM threshold: below it the layer takes the FP32-pipe kernels)
    (5000, 0, 20, 128, 256, 1), (4099, 8, 8, 256, 128, 1), (9000, 1, 3, 512, 64, 1), (4096, 17, 33, 256, 128, 1), (8000, 0, 9, 64, 256, 1),
    (4096 * 2, 4, 4, 128, 128, 2), (5003, 0, 40, 192, 256, 1), (4097, 9, 9, 128, 256, 1), (8192, 2, 6, 512, 64, 2)]
for N, kmin, kmax, fin, fout, G in cases:
    gen = torch.Generator().manual_seed(N + kmax + fin)
    deg = torch.randint(kmin, kmax + 1, (N,), generator=gen); deg[::13] = 0
    tgt = torch.arange(N).repeat_interleave(deg)
    npg = N // G
    src = (torch.randint(0, npg, (int(deg.sum()),), generator=gen) + (tgt // npg) * npg)       # edges stay inside their graph
    ei = torch.stack([src, tgt])
    x = torch.randn(N, fin, generator=gen) * 0.5
    if G > 1: x[npg:] *= 3.0                                                                    # a different logit scale per graph
    x = x.to(torch.bfloat16)
    Ws, As = O.init_gat_params(fin, fout, 4, gen)
    rowptr, col, _ = mg.ops.csr_from_coo(ei.cuda(), N, by_target=True)
    y = mg.ops.gat_forward(x.cuda(), rowptr, col, Ws.cuda(), As.cuda(), concat=False, slope=0.2, out_dtype=torch.float32,
                           nodes_per_graph=(npg if G > 1 else 0))
    torch.cuda.synchronize()
    if G == 1:
        ref = O.gat_layer(x.float(), ei, Ws, As, 0.2, concat=False)
    else:
        ref = torch.cat([O.gat_layer(x[g * npg:(g + 1) * npg].float(), ei[:, (tgt // npg) == g] - g * npg, Ws, As, 0.2, concat=False) for g in range(G)])
    d = (y.cpu() - ref).abs()
    print(N, kmin, kmax, fin, fout, G, "err %.3e zero_rows %.1e nan %d" % (float(d.max()), float(y.cpu()[deg == 0].abs().max()), int(y.isnan().sum())), flush=True)
