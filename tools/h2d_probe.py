#!/usr/bin/env python
"""Where does the end-to-end (host buffers) number stop scaling?  Every rank of a torchrun launch copies a pinned
168 MB buffer (one cfg 2 feature map batch) to its GPU in a loop, first ALONE (ranks take turns), then with 2, 4, ...
ranks copying at the same time; rank 0 prints GB/s per rank and in aggregate.  If the aggregate flattens while the
per-rank rate drops, the limit is the host side (memory / root complexes shared by the GPUs), not the block.
    torchrun --nproc-per-node 8 tools/h2d_probe.py"""
import os, sys, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
nbytes = 16 * 20 * 512 * 512 * 2
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
host.fill_(1)
devbuf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
reps = 30


def copy_rate():
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        devbuf.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


rates = {}
active = 1
while active <= world:
    dist.barrier()
    r = copy_rate() if rank < active else 0.0
    t = torch.tensor([r], device=dev)
    allr = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allr, t)
    rates[active] = [float(v) for v in allr[:active]]
    active *= 2
if rank == 0:
    try:
        aff = sorted(os.sched_getaffinity(0))
        print("rank 0 cpu affinity: %d cpus (%d..%d); host cpus %d" % (len(aff), aff[0], aff[-1], os.cpu_count()))
    except Exception:
        pass
    print("| ranks copying | GB/s per rank (min .. max) | aggregate GB/s |")
    print("|---:|---:|---:|")
    for k, v in rates.items():
        print("| %d | %.1f .. %.1f | %.1f |" % (k, min(v), max(v), sum(v)))
dist.barrier()
dist.destroy_process_group()
