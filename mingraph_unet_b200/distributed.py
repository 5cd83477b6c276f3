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


class InlineGather:
    """NCCL fallback of :class:`PeerExchange` (shapes the fused block kernel does not take, or no peer access): one
    ``all_gather_into_tensor`` per step of the slot's packed small-output buffer, issued from the step's own stream on
    ONE communicator.  ``torch.distributed`` runs a group's collectives on its internal NCCL stream in issue order, so
    the gathers of all slots are totally ordered (concurrent collectives of several communicators hung an 8-GPU run in
    round 1).  Costs the host one collective enqueue per step and every rank one rendezvous per step: 0.129 -> 0.155
    ms/step from 1 to 8 GPUs (SCALE_r01), which is why the peer exchange replaced it as the default."""

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


class PeerExchange:
    """The per-step all-gather of the small per-image outputs WITHOUT a collective call: fused into the block kernel.

    Every rank owns one gathered buffer ``[2 parity][depth slots][world][slice]`` and one flag array ``[depth][world]``
    (``ops.peer_mem_alloc``: cudaMalloc + CUDA IPC handle); the handles are exchanged once over the process group and
    every rank maps every other rank's allocations (NVLink / NVSwitch peer access).  ``slot(i)`` is the descriptor the
    block kernel of pipeline slot ``i`` takes (``GraphBlock(..., _peer=...)`` -> ``mg_block_forward_push``): while it
    computes, the kernel stores ``l_partition | region_features | hard_labels`` of its images into slice ``rank`` of every
    rank's buffer and its last CTA publishes the slot's step number in every rank's flag array.  No kernel of the step
    waits for another GPU, so a slow rank never makes a fast one hold SMs (what concurrent NCCL kernels did at 8 GPUs).

    Consumer: ``wait(slot)`` enqueues a one-CTA kernel on the current stream that returns once every rank's flag has
    reached this rank's own step count of the slot (bounded: ``status`` turns 1 instead of hanging); ``views(slot)`` are
    the gathered tensors of the slot's latest step.  Consumers must be stream-ordered before the slot's next step (the
    parity double buffer then makes acknowledgements unnecessary, csrc/peer_exchange.cu).  ``epilogues()`` are
    ``wait`` closures to record at the end of each slot's CUDA graph.

    Call ``reset()`` (collective) after the step graphs were built: their warm-up passes push as well.  ``close()``
    (collective) unmaps and frees."""

    def __init__(self, B: int, N: int, K: int, D: int, device, depth: int, group: Optional[dist.ProcessGroup] = None,
                 _local=None):
        from . import ops
        self._ops = ops
        self._group = group
        self.B, self.N, self.K, self.D, self.depth = B, N, K, D, depth
        self.device = torch.device(device)
        self.n_small = B * (1 + K * D + N)
        self.n_pad = (self.n_small + 31) // 32 * 32                   # 128-byte slices
        self.packed = [torch.zeros(self.n_small, dtype=torch.float32, device=device) for _ in range(depth)]
        if _local is not None:                                        # local_group(): all "ranks" in this process
            self.rank, self.world, bufs, flags = _local
            self._own, self._opened = None, []
        else:
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
            buf_bytes, flag_bytes = self._sizes(self.world)
            self._own = [ops.peer_mem_alloc(buf_bytes, device), ops.peer_mem_alloc(flag_bytes, device)]
            handles = [None] * self.world
            dist.all_gather_object(handles, (self._own[0][1], self._own[1][1]), group=group)
            self._opened = []
            bufs, flags = [], []
            for r in range(self.world):
                if r == self.rank:
                    bufs.append(self._own[0][0])
                    flags.append(self._own[1][0])
                else:
                    pb, pf = ops.peer_mem_open(handles[r][0], device), ops.peer_mem_open(handles[r][1], device)
                    self._opened += [pb, pf]
                    bufs.append(pb)
                    flags.append(pf)
        self._map(bufs, flags)
        if _local is None:
            torch.cuda.synchronize(self.device)
            dist.barrier(group)                                       # every mapping exists before anybody pushes

    def _sizes(self, world: int):
        return 2 * self.depth * world * self.n_pad * 4, max(256, self.depth * world * 4)

    def _map(self, bufs, flags) -> None:
        from . import _lib
        ops, W, depth, device = self._ops, self.world, self.depth, self.device
        buf_bytes, flag_bytes = self._sizes(W)
        self._gathered = ops.tensor_from_ptr(bufs[self.rank], buf_bytes, device).view(torch.float32).view(2, depth, W, self.n_pad)
        self._flags = ops.tensor_from_ptr(flags[self.rank], flag_bytes, device).view(torch.int32)
        self._ptr_bufs = torch.tensor(bufs, dtype=torch.int64, device=device)
        self._ptr_flags = torch.tensor(flags, dtype=torch.int64, device=device)
        self._seq = torch.zeros(depth, dtype=torch.int32, device=device)
        self._done = torch.zeros(depth, dtype=torch.int32, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        self._slots = []
        for i in range(depth):
            d = _lib.PeerOut()
            d.peer_bufs_dev, d.peer_flags_dev = self._ptr_bufs.data_ptr(), self._ptr_flags.data_ptr()
            d.world, d.rank = W, self.rank
            d.slice_offset = (i * W + self.rank) * self.n_pad
            d.parity_stride = depth * W * self.n_pad
            d.flag_index = i * W + self.rank
            d.seq, d.done = self._seq[i:].data_ptr(), self._done[i:].data_ptr()
            self._slots.append(d)
        self._steps = [0] * depth                                     # host mirror of seq (steps since reset)

    @classmethod
    def local_group(cls, world: int, B: int, N: int, K: int, D: int, device, depth: int):
        """``world`` exchange endpoints inside ONE process on ONE device (tests, single-GPU checks): the same kernels
        and index arithmetic, the "peers" being plain allocations of the same GPU.  Steps of different endpoints must
        be enqueued before any ``wait`` that needs them (one GPU cannot be relied on to run kernels that wait for each
        other concurrently)."""
        from . import ops
        probe = cls.__new__(cls)
        probe.depth, probe.n_pad = depth, (B * (1 + K * D + N) + 31) // 32 * 32
        buf_bytes, flag_bytes = cls._sizes(probe, world)
        allocs = [(ops.peer_mem_alloc(buf_bytes, device)[0], ops.peer_mem_alloc(flag_bytes, device)[0]) for _ in range(world)]
        bufs, flags = [a[0] for a in allocs], [a[1] for a in allocs]
        group = [cls(B, N, K, D, device, depth, _local=(r, world, bufs, flags)) for r in range(world)]
        group[0]._local_allocs = bufs + flags                         # freed by group[0].close()
        return group

    def slot(self, i: int):
        return self._slots[i]

    def slots(self):
        return list(self._slots)

    def reset(self) -> None:
        """Collective.  Zero the step counters and flags once all ranks are idle (after the graphs were built)."""
        torch.cuda.synchronize(self.device)
        self._barrier()                                               # every rank's warm-up pushes have landed
        self._flags.zero_()
        self._seq.zero_()
        self._done.zero_()
        self.status.zero_()
        self._steps = [0] * self.depth
        torch.cuda.synchronize(self.device)
        self._barrier()                                               # nobody pushes before every flag is zeroed

    def _barrier(self) -> None:
        if self._own is not None:
            dist.barrier(self._group)

    def stepped(self, slot: int) -> None:
        """Tell the host mirror that one step of ``slot`` was enqueued (``views`` needs its parity)."""
        self._steps[slot] += 1

    def wait(self, slot: int) -> None:
        """Current stream waits for every rank's payload of the slot's latest enqueued step."""
        self._ops.peer_wait(self._flags, slot * self.world, self.world, self._seq[slot:slot + 1], self.status)

    def epilogues(self):
        return [(lambda _outputs, i=i: self.wait(i)) for i in range(self.depth)]

    def views(self, slot: int) -> GatheredOutputs:
        """Gathered outputs of the slot's latest step (valid on the stream that ran ``wait(slot)``)."""
        if self._steps[slot] < 1:
            raise RuntimeError("PeerExchange.views() before the slot's first step (call stepped(slot) per submit)")
        B, K, D, N, W = self.B, self.K, self.D, self.N, self.world
        g = self._gathered[(self._steps[slot] - 1) & 1, slot]
        loss = g[:, :B].reshape(W * B)
        reg = g[:, B:B + B * K * D].reshape(W * B, K, D)
        lab = g[:, B + B * K * D:self.n_small].view(torch.int32).reshape(W * B, N)
        return GatheredOutputs(loss, reg, lab)

    def close(self) -> None:
        """Collective.  Unmap the peers' allocations, then free the own ones."""
        torch.cuda.synchronize(self.device)
        self._gathered = self._flags = None
        if self._own is None:                                         # local_group(): endpoint 0 owns every allocation
            for ptr in getattr(self, "_local_allocs", []):
                self._ops.peer_mem_free(ptr)
            self._local_allocs = []
            return
        dist.barrier(self._group)
        for p in self._opened:
            self._ops.peer_mem_close(p)
        self._opened = []
        dist.barrier(self._group)                                     # nobody maps the buffers any more
        for ptr, _ in self._own:
            self._ops.peer_mem_free(ptr)
        self._own = None
