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
