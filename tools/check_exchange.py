#!/usr/bin/env python
"""Check of the multi-GPU exchange (distributed.PeerExchange fused into the block kernel, or the NCCL fallback
InlineGather; --mode p2p | inline) at any world size: every rank pipelines several steps WITHOUT any host-side flow
control between ranks, then verifies that each step's GATHERED per-image outputs (loss, region features, labels of ALL
ranks) equal what the eager block computes for every rank's input of that step (inputs are seeded by (step, rank), so
each rank can recompute the others').  Odd ranks are slowed down with device-side sleeps so that ranks drift apart by
more than a step: the parity double buffer of the peer exchange has to keep the payloads intact.
Launch: torchrun --nproc-per-node N tools/check_exchange.py [--mode p2p], or plain `python` for a 1-rank group."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
import mingraph_unet_b200 as mg
from mingraph_unet_b200.distributed import InlineGather, PeerExchange

MODE = "p2p"
for i, a in enumerate(sys.argv):
    if a.startswith("--mode="):
        MODE = a.split("=", 1)[1]
    elif a == "--mode" and i + 1 < len(sys.argv):
        MODE = sys.argv[i + 1]
assert MODE in ("inline", "p2p"), MODE

B, C, H, W, D, K, depth, steps = 4, 20, 128, 96, 64, 2, 3, 14
N = (H // 16) * (W // 16)
torch.manual_seed(1234)
blk = mg.GraphBlock(node_feature_dim=C, num_segments=K).to(dev).eval()


def make_input(step, r):
    return torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(1000 * step + r)).to(dev)


if MODE == "p2p":
    ex = PeerExchange(B, N, K, D, dev, depth)
    pipe = mg.PipelinedGraphBlock(blk, make_input(0, rank), image_size=(H, W), depth=depth, packed_small=ex.packed,
                                  peers=ex.slots(), epilogues=ex.epilogues())
    ex.reset()                          # the graphs' warm-up passes pushed too: restart the step counters together
else:
    ex = InlineGather(B, N, K, D, dev, depth)
    pipe = mg.PipelinedGraphBlock(blk, make_input(0, rank), image_size=(H, W), depth=depth, packed_small=ex.packed)
got = []
for s in range(steps):
    if rank % 2 == 1 and s % 3 == 0:
        torch.cuda._sleep(2_000_000)        # ~1 ms: this rank falls several steps behind the even ones
    slot, out = pipe.submit(make_input(s, rank))
    if MODE == "inline":
        ex.gather(slot, pipe.stream(slot))        # one collective enqueue from the step's own stream
    else:
        ex.stepped(slot)                          # (the wait for every rank's payload is the graph's last node)
    with torch.cuda.stream(pipe.stream(slot)):
        g = ex.views(slot)
        got.append((g.l_partition.clone(), g.region_features.clone(), g.hard_labels.clone(), out.l_partition.clone()))
    pipe.mark(slot)
pipe.join()
torch.cuda.synchronize()
ok = True
with torch.no_grad():
    for s, (loss, reg, lab, own_loss) in enumerate(got):
        for r in range(world):
            ref = blk(feature_map=make_input(s, r), image_size=(H, W), want_dense=False)
            sl = slice(r * B, (r + 1) * B)
            ok &= torch.equal(loss[sl], ref.l_partition) and torch.equal(reg[sl], ref.region_features)
            ok &= torch.equal(lab[sl], ref.hard_labels)
        ok &= torch.equal(own_loss, loss[rank * B:(rank + 1) * B])
if MODE == "p2p":
    ok &= int(ex.status.item()) == 0        # no wait ran into its spin bound
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("%s exchange world=%d: %s" % (MODE, world, "OK" if int(flag) else "MISMATCH"), flush=True)
code = 0 if int(flag) else 1
del pipe
if MODE == "p2p":
    ex.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(code)
