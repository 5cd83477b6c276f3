"""Multi-GPU plumbing: one process per GPU, the batch sharded by image.

Every image's graph is independent (scripts/train_end_to_end.py:300-425 builds one graph per
image), so there are no cross-GPU edges and no data-path collective inside the block.  The only
exchanges are (i) an all-gather of the SMALL per-image outputs — N-cut loss, region features and
patch labels, from which the node-level ``(B,N,D)`` output is ``region[labels]`` — and (ii) in
training an all-reduce of the graph-layer gradients (22 792 fp32 = 91 KB for the default block).
The dense ``(B,D,H,W)`` map is consumed locally by the fusion stage and is never gathered.
"""
from __future__ import annotations

from typing import List, NamedTuple, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous image range ``[lo, hi)`` of ``rank``; the first ``global_batch % world`` ranks
    take one extra image."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def bind_to_gpu_numa(device) -> str:
    """Bind the calling thread (and the threads it creates later) to the CPUs closest to ``device`` with
    ``nvmlDeviceSetCpuAffinity``.  With one process per GPU this keeps each rank's pinned host buffers (first-touch) and
    its launch thread on the GPU's own NUMA node, so the ranks' host<->device copies do not all cross one socket's
    memory controllers.  Call it BEFORE allocating pinned memory.  Returns a short description of what happened (never
    raises: affinity is an optimisation).  Written after the round-1 GPU budget was spent — effect not yet measured."""
    try:
        import pynvml
        pynvml.nvmlInit()
        dev = torch.device(device)
        handle = None
        try:
            pr = torch.cuda.get_device_properties(dev)
            bus_id = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(dev.index if dev.index is not None else 0)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        import os
        n = len(os.sched_getaffinity(0))
        return f"bound to the {n} CPUs nearest the GPU (nvmlDeviceSetCpuAffinity)"
    except Exception as e:          # pragma: no cover - depends on the box
        return f"not bound ({type(e).__name__}: {e})"


class GatheredOutputs(NamedTuple):
    l_partition: torch.Tensor        # (Bg,)
    region_features: torch.Tensor    # (Bg, K, D)
    hard_labels: torch.Tensor        # (Bg, N) int32

    def node_features(self) -> torch.Tensor:
        """Node-level output ``(Bg, N, D) = region_features[labels]`` (train_end_to_end.py:403-406)."""
        idx = self.hard_labels.long().unsqueeze(-1).expand(-1, -1, self.region_features.shape[-1])
        return torch.gather(self.region_features, 1, idx)


def gather_block_outputs(l_partition: torch.Tensor, region_features: torch.Tensor, hard_labels: torch.Tensor,
                         global_batch: int, group: Optional[dist.ProcessGroup] = None) -> GatheredOutputs:
    """All-gather the per-image outputs of every rank, in image order.  Shards may differ in size by
    one image (``shard_range``); ranks pad to the largest shard so a single fixed-size collective
    per dtype is issued."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_range(global_batch, rank, world)
    B = hi - lo
    if l_partition.shape[0] != B or region_features.shape[0] != B or hard_labels.shape[0] != B:
        raise ValueError(f"rank {rank} holds {l_partition.shape[0]} images, expected {B}")
    Bmax = -(-global_batch // world)
    K, D = region_features.shape[1:]
    N = hard_labels.shape[1]
    dev = l_partition.device
    f = torch.zeros(Bmax, 1 + K * D, dtype=torch.float32, device=dev)
    f[:B, 0] = l_partition.float()
    f[:B, 1:] = region_features.reshape(B, K * D).float()
    i = torch.zeros(Bmax, N, dtype=torch.int32, device=dev)
    i[:B] = hard_labels.to(torch.int32)
    gf = torch.empty(world * Bmax, 1 + K * D, dtype=torch.float32, device=dev)
    gi = torch.empty(world * Bmax, N, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(gf, f, group=group)
    dist.all_gather_into_tensor(gi, i, group=group)
    keep: List[int] = []
    for r in range(world):
        rlo, rhi = shard_range(global_batch, r, world)
        keep.extend(range(r * Bmax, r * Bmax + (rhi - rlo)))
    sel = torch.tensor(keep, dtype=torch.long, device=dev)
    gf, gi = gf.index_select(0, sel), gi.index_select(0, sel)
    return GatheredOutputs(gf[:, 0].contiguous(), gf[:, 1:].reshape(global_batch, K, D).contiguous(), gi.contiguous())


def allreduce_graph_grads(module: torch.nn.Module, group: Optional[dist.ProcessGroup] = None) -> int:
    """Average the gradients of ``module``'s parameters over ranks with ONE flat all-reduce
    (latency-bound: the default block has 22 792 parameters).  Returns the number of elements."""
    params = [p for p in module.parameters() if p.grad is not None]
    if not params:
        return 0
    flat = torch.cat([p.grad.reshape(-1).float() for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return off


class OverlappedGather:
    """Per-step all-gather of the small per-image outputs, issued on a side stream so that it overlaps
    the next step's kernels instead of extending the step by the collective's launch latency.

    ``push(out)`` packs ``(l_partition | region_features | hard_labels)`` of the local shard into one of
    two rotating buffers on the current stream and enqueues ONE ``all_gather_into_tensor`` on the side
    stream; ``latest()`` makes the current stream wait for the most recent gather and returns
    :class:`GatheredOutputs` views of its result.  With two buffers a step never waits for its own
    gather, only (at most) for the one issued two steps earlier.  Equal shard sizes are required
    (``global_batch % world == 0``); use :func:`gather_block_outputs` otherwise."""

    def __init__(self, B: int, N: int, K: int, D: int, device, group: Optional[dist.ProcessGroup] = None):
        self.B, self.N, self.K, self.D, self.group = B, N, K, D, group
        self.world = dist.get_world_size(group)
        self.n_small = B * (1 + K * D + N)
        self.packed = [torch.empty(self.n_small, dtype=torch.float32, device=device) for _ in range(2)]
        self.gathered = [torch.empty(self.world * self.n_small, dtype=torch.float32, device=device) for _ in range(2)]
        self.done = [None, None]
        self.side = torch.cuda.Stream(device=device)
        self.turn = 0

    def push(self, l_partition: torch.Tensor, region_features: torch.Tensor, hard_labels: torch.Tensor) -> None:
        i = self.turn
        cur = torch.cuda.current_stream()
        if self.done[i] is not None:
            cur.wait_event(self.done[i])              # the gather that last read this buffer (two steps ago)
        B, K, D, N = self.B, self.K, self.D, self.N
        buf = self.packed[i]
        torch.cat([l_partition.reshape(-1), region_features.reshape(-1), hard_labels.reshape(-1).view(torch.float32)],
                  out=buf)
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            dist.all_gather_into_tensor(self.gathered[i], buf, group=self.group)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.done[i] = ev
        self.turn = 1 - i

    def latest(self) -> GatheredOutputs:
        i = 1 - self.turn
        if self.done[i] is None:
            raise RuntimeError("OverlappedGather.latest() before any push()")
        torch.cuda.current_stream().wait_event(self.done[i])
        B, K, D, N, W = self.B, self.K, self.D, self.N, self.world
        g = self.gathered[i].view(W, self.n_small)
        loss = g[:, :B].reshape(W * B)
        reg = g[:, B:B + B * K * D].reshape(W * B, K, D)
        lab = g[:, B + B * K * D:].reshape(W * B, N).view(torch.int32)
        return GatheredOutputs(loss, reg, lab)

    def drain(self) -> None:
        for ev in self.done:
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)


class CapturedGather:
    """The per-step all-gather of the small per-image outputs recorded INSIDE each pipeline slot's CUDA graph.

    ``OverlappedGather`` costs the host a pack kernel launch, two event operations and one NCCL enqueue per step; at
    cfg 2 a pipelined step is ~130 us of GPU time, so that host work (not the 74 KB transfer) set the multi-GPU step
    time.  Here the block kernel writes ``l_partition | region_features | hard_labels`` straight into the slot's
    packed buffer (``CapturedGraphBlock(packed_small=...)``: the outputs are views of it, no pack kernel) and the graph
    ends with ONE ``all_gather_into_tensor`` of that buffer, so a step stays a single ``cudaGraphLaunch`` on every
    rank.  Each slot has its own communicator: graphs of different slots replay concurrently on different streams, and
    collectives of ONE communicator must not be issued concurrently."""

    def __init__(self, B: int, N: int, K: int, D: int, device, depth: int):
        self.B, self.N, self.K, self.D, self.depth = B, N, K, D, depth
        self.world = dist.get_world_size()
        self.n_small = B * (1 + K * D + N)
        self.packed = [torch.zeros(self.n_small, dtype=torch.float32, device=device) for _ in range(depth)]
        self.gathered = [torch.zeros(self.world * self.n_small, dtype=torch.float32, device=device) for _ in range(depth)]
        self.groups = [dist.new_group(backend="nccl") for _ in range(depth)]        # collective: same order on all ranks

    def epilogue(self, slot: int):
        def run(_outputs) -> None:
            dist.all_gather_into_tensor(self.gathered[slot], self.packed[slot], group=self.groups[slot])
        return run

    def epilogues(self):
        return [self.epilogue(i) for i in range(self.depth)]

    def views(self, slot: int) -> GatheredOutputs:
        """The slot's gathered result (valid on the slot's stream after its replay)."""
        B, K, D, N, W = self.B, self.K, self.D, self.N, self.world
        g = self.gathered[slot].view(W, self.n_small)
        loss = g[:, :B].reshape(W * B)
        reg = g[:, B:B + B * K * D].reshape(W * B, K, D)
        lab = g[:, B + B * K * D:].reshape(W * B, N).view(torch.int32)
        return GatheredOutputs(loss, reg, lab)


class InlineGather:
    """The per-step all-gather issued straight from the step's own stream, on ONE communicator, with no pack kernel:
    ``gather(slot, stream)`` all-gathers the slot's packed small-output buffer (the block kernel wrote
    ``l_partition | region_features | hard_labels`` into it in place, ``CapturedGraphBlock(packed_small=...)``).

    ``torch.distributed`` runs every collective of a process group on the group's internal NCCL stream, in issue order:
    that stream waits for the caller's stream (the step's replay), and the caller's stream — whose next work is the
    slot's replay ``depth`` steps later — waits for the collective.  So the gathers of all slots are serialised on one
    communicator in step order (no concurrent collectives, unlike :class:`CapturedGather`), nothing blocks the other
    slots, and the host pays one collective enqueue per step (``OverlappedGather`` additionally launches a pack kernel
    and four event operations)."""

    def __init__(self, B: int, N: int, K: int, D: int, device, depth: int, group: Optional[dist.ProcessGroup] = None):
        self.B, self.N, self.K, self.D, self.depth, self.group = B, N, K, D, depth, group
        self.world = dist.get_world_size(group)
        self.n_small = B * (1 + K * D + N)
        self.packed = [torch.zeros(self.n_small, dtype=torch.float32, device=device) for _ in range(depth)]
        self.gathered = [torch.zeros(self.world * self.n_small, dtype=torch.float32, device=device) for _ in range(depth)]

    def gather(self, slot: int, stream=None) -> None:
        if stream is None:
            dist.all_gather_into_tensor(self.gathered[slot], self.packed[slot], group=self.group)
            return
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(self.gathered[slot], self.packed[slot], group=self.group)

    def views(self, slot: int) -> GatheredOutputs:
        """The slot's gathered result (valid on the stream ``gather`` was issued from)."""
        B, K, D, N, W = self.B, self.K, self.D, self.N, self.world
        g = self.gathered[slot].view(W, self.n_small)
        loss = g[:, :B].reshape(W * B)
        reg = g[:, B:B + B * K * D].reshape(W * B, K, D)
        lab = g[:, B + B * K * D:].reshape(W * B, N).view(torch.int32)
        return GatheredOutputs(loss, reg, lab)


class BucketedGather:
    """The exchange of several consecutive steps in ONE collective (opt-in, ``bench.py --exchange bucketed``; written
    after the round-1 GPU budget was spent — host logic covered by the gloo world-2 test, not yet timed on hardware).

    The per-step payload is 74 KB at cfg 2 while a pipelined step is ~125 us of GPU time: a per-step collective costs
    the host one NCCL enqueue and every rank one rendezvous per step.  Here the ``depth`` pipeline slots' packed
    buffers are consecutive views of ONE ring; the ring is cut into buckets of ``bucket`` consecutive slots and a
    bucket is gathered with one ``all_gather_into_tensor`` on a side stream after its LAST slot's step was enqueued
    (the side stream waits for the events of exactly the bucket's steps).  A slot's next replay (``depth`` steps later)
    waits for the gather that read its buffer, so with ``depth >= 2 * bucket`` that gather has the other buckets' steps
    as slack.  One communicator, collectives in step order (same liveness argument as :class:`InlineGather`); results
    of a step become visible up to ``bucket - 1`` steps late — a throughput, not a latency, optimisation.

    Use: ``before(slot, stream)`` ahead of the slot's submit, ``after(slot, stream)`` behind it; ``flush()`` gathers a
    partly filled bucket (end of a run); ``views(slot)`` as for the other variants."""

    def __init__(self, B: int, N: int, K: int, D: int, device, depth: int, bucket: Optional[int] = None,
                 group: Optional[dist.ProcessGroup] = None):
        if bucket is None:
            bucket = max(1, depth // 2)
        if bucket < 1 or depth % bucket != 0:
            raise ValueError("depth must be a multiple of bucket")
        self.B, self.N, self.K, self.D, self.depth, self.bucket, self.group = B, N, K, D, depth, bucket, group
        self.world = dist.get_world_size(group)
        self.n_small = B * (1 + K * D + N)
        self._cuda = torch.device(device).type == "cuda"
        self._ring = torch.zeros(depth * self.n_small, dtype=torch.float32, device=device)
        self.packed = [self._ring[i * self.n_small:(i + 1) * self.n_small] for i in range(depth)]
        nb = depth // bucket
        self.gathered = [torch.zeros(self.world * bucket * self.n_small, dtype=torch.float32, device=device)
                         for _ in range(nb)]
        self.side = torch.cuda.Stream(device=device) if self._cuda else None
        self._ready = [torch.cuda.Event() for _ in range(depth)] if self._cuda else None
        self._done = [torch.cuda.Event() for _ in range(nb)] if self._cuda else None
        self._done_valid = [False] * nb
        self._pending: List[int] = []            # slots stepped since the last gather, in step order

    def before(self, slot: int, stream=None) -> None:
        """The slot's stream waits for the gather that last read the slot's packed buffer."""
        b = slot // self.bucket
        if self._cuda and self._done_valid[b]:
            stream.wait_event(self._done[b])

    def after(self, slot: int, stream=None) -> None:
        """Call after the slot's step was enqueued on ``stream``; gathers the bucket when this was its last slot."""
        if self._pending and (slot != self._pending[-1] + 1 or slot // self.bucket != self._pending[0] // self.bucket):
            raise RuntimeError("BucketedGather: slots must be stepped round-robin (0, 1, ..., depth-1, 0, ...)")
        if self._cuda:
            self._ready[slot].record(stream)
        self._pending.append(slot)
        if (slot + 1) % self.bucket == 0:
            self._gather()

    def flush(self) -> None:
        """Gather a partly filled bucket (collective: every rank must call it at the same step)."""
        if self._pending:
            self._gather()

    def _gather(self) -> None:
        b = self._pending[0] // self.bucket
        src = self._ring[b * self.bucket * self.n_small:(b + 1) * self.bucket * self.n_small]
        if self._cuda:
            for s in self._pending:
                self.side.wait_event(self._ready[s])
            with torch.cuda.stream(self.side):
                dist.all_gather_into_tensor(self.gathered[b], src, group=self.group)
                self._done[b].record(self.side)
            self._done_valid[b] = True
        else:
            dist.all_gather_into_tensor(self.gathered[b], src, group=self.group)
        self._pending = []

    def drain(self) -> None:
        """Current stream waits for every gather issued so far."""
        if self._cuda:
            cur = torch.cuda.current_stream()
            for ok, ev in zip(self._done_valid, self._done):
                if ok:
                    cur.wait_event(ev)

    def views(self, slot: int) -> GatheredOutputs:
        """The gathered result of the slot's most recent GATHERED step (valid after ``drain()`` / on the side stream)."""
        B, K, D, N, W = self.B, self.K, self.D, self.N, self.world
        g = self.gathered[slot // self.bucket].view(W, self.bucket, self.n_small)[:, slot % self.bucket]
        loss = g[:, :B].reshape(W * B)
        reg = g[:, B:B + B * K * D].reshape(W * B, K, D)
        lab = g[:, B + B * K * D:].view(torch.int32).reshape(W * B, N)
        return GatheredOutputs(loss, reg, lab)


class PeerGather:
    """EXPERIMENTAL — compiled and wired, not yet run on hardware (the round-1 GPU budget was spent); opt-in via
    ``bench.py --exchange p2p`` / ``tools/check_captured_gather.py --mode p2p``.

    The per-step exchange of the small outputs as plain NVLink stores (``csrc/peer_push.cu``) instead of a collective:
    every rank owns a symmetric gathered buffer (``depth`` slots x ``world`` slices) mapped into all peers through
    ``torch.distributed._symmetric_memory``; the epilogue of a step's graph is ONE kernel that stores the slot's packed
    payload into slice ``rank`` of every peer's buffer and publishes a sequence number.  No kernel of this scheme waits
    for another GPU, so a slow rank never makes a fast rank hold SMs (the failure mode of concurrent NCCL kernels).
    Call ``reset()`` once after the step graphs were built (their warm-up passes push too).
    Consumers call ``wait(slot)`` exactly once per step before reading ``views(slot)``.  Flow control is the caller's:
    a producer may overwrite slot ``s`` of a peer ``depth`` steps later, so a consumer that reads the gathered data must
    keep ranks within ``depth`` steps of each other (the bench does not read it; its end-of-region barrier closes it)."""

    def __init__(self, B: int, N: int, K: int, D: int, device, depth: int, group: Optional[dist.ProcessGroup] = None):
        import torch.distributed._symmetric_memory as symm
        from . import ops
        self._ops = ops
        group = group if group is not None else dist.group.WORLD
        self._group = group
        self.B, self.N, self.K, self.D, self.depth = B, N, K, D, depth
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n_small = B * (1 + K * D + N)
        self.n_pad = (self.n_small + 3) // 4 * 4                      # 16-byte payloads
        self._packed = [torch.zeros(self.n_pad, dtype=torch.float32, device=device) for _ in range(depth)]
        self.packed = [p[:self.n_small] for p in self._packed]       # what the block kernel writes (views)
        self._gathered = symm.empty(depth * self.world * self.n_pad, dtype=torch.float32, device=device)
        self._flags = symm.empty(depth * self.world, dtype=torch.int32, device=device)
        self._gathered.zero_()
        self._flags.zero_()
        self._h_buf = symm.rendezvous(self._gathered, group)
        self._h_flag = symm.rendezvous(self._flags, group)
        self._seq = torch.zeros(depth, self.world, dtype=torch.int32, device=device)
        self._wseq = torch.zeros(depth, self.world, dtype=torch.int32, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)                                            # nobody pushes before every flag is zeroed

    def reset(self) -> None:
        """Collective.  Zero every sequence counter and flag once all ranks are idle.  Must be called after the graphs
        that contain the push were built: building a ``CapturedGraphBlock`` runs its epilogue eagerly ``warmup`` times,
        so without this the producers' counters would start ahead of the consumers' and the first ``wait`` of a slot
        would be satisfied by a warm-up push (stale payload)."""
        torch.cuda.synchronize(self._flags.device)
        dist.barrier(self._group)                                      # every rank's warm-up pushes have landed
        self._flags.zero_()
        self._seq.zero_()
        self._wseq.zero_()
        self.status.zero_()
        torch.cuda.synchronize(self._flags.device)
        dist.barrier(self._group)                                      # nobody pushes before every flag is zeroed

    def epilogue(self, slot: int):
        off = (slot * self.world + self.rank) * self.n_pad * 4
        flag = slot * self.world + self.rank

        def run(_outputs) -> None:
            self._ops.peer_push(self._packed[slot], self._h_buf.buffer_ptrs_dev, self.world, off,
                                self._h_flag.buffer_ptrs_dev, flag, self._seq[slot])
        return run

    def epilogues(self):
        return [self.epilogue(i) for i in range(self.depth)]

    def wait(self, slot: int) -> None:
        """Current stream waits for every rank's payload of this slot's next step (call once per step)."""
        self._ops.peer_wait(self._flags, slot * self.world, self.world, self._wseq[slot], self.status)

    def views(self, slot: int) -> GatheredOutputs:
        B, K, D, N, W = self.B, self.K, self.D, self.N, self.world
        g = self._gathered[slot * W * self.n_pad:(slot + 1) * W * self.n_pad].view(W, self.n_pad)
        loss = g[:, :B].reshape(W * B)
        reg = g[:, B:B + B * K * D].reshape(W * B, K, D)
        lab = g[:, B + B * K * D:self.n_small].reshape(W * B, N).view(torch.int32)
        return GatheredOutputs(loss, reg, lab)
