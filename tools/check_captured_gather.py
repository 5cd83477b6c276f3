#!/usr/bin/env python
"""Check of the multi-GPU exchange variants (distributed.CapturedGather / InlineGather / PeerGather / BucketedGather +
PipelinedGraphBlock; --mode captured | inline | p2p | bucketed) at any world size:
every rank pipelines several steps, then verifies that each step's GATHERED per-image outputs (loss, region features,
labels of ALL ranks) equal what the eager block computes for every rank's input of that step (inputs are seeded by
(step, rank), so each rank can recompute the others').  Launch: torchrun --nproc-per-node N tools/check_captured_gather.py,
or plain `python` with RANK/WORLD_SIZE/MASTER_* unset for a 1-rank NCCL group."""
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
from mingraph_unet_b200.distributed import BucketedGather, CapturedGather, InlineGather, PeerGather

MODE = "captured"                       # --mode captured | inline | p2p | bucketed
for i, a in enumerate(sys.argv):
    if a.startswith("--mode="):
        MODE = a.split("=", 1)[1]
    elif a == "--mode" and i + 1 < len(sys.argv):
        MODE = sys.argv[i + 1]
assert MODE in ("captured", "inline", "p2p", "bucketed"), MODE

B, C, H, W, D, K, depth, steps = 4, 20, 128, 96, 64, 2, 2, 5
if MODE == "bucketed":
    depth, steps = 4, 11                # buckets of two slots; the last step leaves a partly filled bucket
N = (H // 16) * (W // 16)
torch.manual_seed(1234)
blk = mg.GraphBlock(node_feature_dim=C, num_segments=K).to(dev).eval()


def make_input(step, r):
    return torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(1000 * step + r)).to(dev)


gather = {"captured": CapturedGather, "inline": InlineGather, "p2p": PeerGather,
          "bucketed": BucketedGather}[MODE](B, N, K, D, dev, depth)
pipe = mg.PipelinedGraphBlock(blk, make_input(0, rank), image_size=(H, W), depth=depth, packed_small=gather.packed,
                              epilogues=None if MODE in ("inline", "bucketed") else gather.epilogues(),
                              epilogue_parallel=MODE == "p2p")
if MODE == "p2p":
    gather.reset()                      # the graphs' warm-up passes pushed too: restart the sequence numbers together
got = []


def collect_bucket(step_slots, own):
    """bucketed: the gathers issued so far are complete on the current stream; copy the results of these steps out."""
    gather.drain()
    for slot_, own_loss in zip(step_slots, own):
        g = gather.views(slot_)
        got.append((g.l_partition.clone(), g.region_features.clone(), g.hard_labels.clone(), own_loss))
    gather.side.wait_stream(torch.cuda.current_stream())      # the next gather into this buffer comes after the copies


open_slots, open_own = [], []
for s in range(steps):
    if MODE == "bucketed":
        nxt = pipe.next_slot
        gather.before(nxt, pipe.stream(nxt))
        slot, out = pipe.submit(make_input(s, rank))
        with torch.cuda.stream(pipe.stream(slot)):
            open_own.append(out.l_partition.clone())
        pipe.mark(slot)
        open_slots.append(slot)
        gather.after(slot, pipe.stream(slot))
        last = s == steps - 1
        if last:
            gather.flush()
        if (slot + 1) % gather.bucket == 0 or last:
            collect_bucket(open_slots, open_own)
            open_slots, open_own = [], []
        continue
    slot, out = pipe.submit(make_input(s, rank))
    if MODE == "inline":
        gather.gather(slot, pipe.stream(slot))        # one collective enqueue from the step's own stream
    with torch.cuda.stream(pipe.stream(slot)):
        if MODE == "p2p":
            gather.wait(slot)                 # every rank's payload of this step has landed
        g = gather.views(slot)
        got.append((g.l_partition.clone(), g.region_features.clone(), g.hard_labels.clone(), out.l_partition.clone()))
    pipe.mark(slot)
    if MODE == "p2p" and (s + 1) % depth == 0:
        pipe.join()
        torch.cuda.synchronize()
        dist.barrier()                        # flow control: no rank overwrites a slot a peer has not read yet
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
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("%s gather world=%d: %s" % (MODE, world, "OK" if int(flag) else "MISMATCH"), flush=True)
# the slots' graphs hold NCCL kernels: leave without tearing the communicators down under them
code = 0 if int(flag) else 1
dist.barrier()
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(code)
