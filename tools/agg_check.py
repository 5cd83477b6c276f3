"""One aggregation parity case in its own process (a hang then costs one short timeout, not a test session):
python tools/agg_check.py N kmin kmax F concat(0/1) out(bf16|f32)"""
import sys, torch
sys.path.insert(0, ".")
from oracle import restate as O
import mingraph_unet_b200 as mg
N, kmin, kmax, F, concat = (int(v) for v in sys.argv[1:6]); out_dtype = torch.bfloat16 if sys.argv[6] == "bf16" else torch.float32
gen = torch.Generator().manual_seed(N + kmax)
deg = torch.randint(kmin, kmax + 1, (N,), generator=gen); deg[::13] = 0
tgt = torch.arange(N).repeat_interleave(deg); src = torch.randint(0, N, (int(deg.sum()),), generator=gen)
ei = torch.stack([src, tgt])
x = torch.randn(N, 64, generator=gen)
if concat: x *= 0.5
x = x.to(torch.bfloat16)
Ws, As = O.init_gat_params(64, F, 4, gen)
rowptr, col, _ = mg.ops.csr_from_coo(ei.cuda(), N, by_target=True)
torch.cuda.synchronize(); print("csr ok; launch", flush=True)
y = mg.ops.gat_forward(x.cuda(), rowptr, col, Ws.cuda(), As.cuda(), concat=bool(concat), slope=0.2, out_dtype=out_dtype)
torch.cuda.synchronize()
print("done", flush=True)
ref = O.gat_layer(x.float(), ei, Ws, As, 0.2, concat=bool(concat))
print("err %.3e zero_rows %.1e" % (float((y.float().cpu() - ref).abs().max()), float(y.float().cpu()[deg == 0].abs().max())), flush=True)
