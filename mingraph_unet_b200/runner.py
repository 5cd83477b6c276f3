"""CUDA-graph replay of the graph block for a fixed input shape.

The block is a short chain of small kernels around two HBM-streaming ones (pool, un-pool); at the
image sizes of the reference's configs the host-side launch path costs more than the kernels.
``CapturedGraphBlock`` records one forward (pool -> fused block -> un-pool into the caller's fusion
buffer) into a CUDA graph and replays it with a single driver call per step.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .block import GraphBlock, GraphBlockOutput


class CapturedGraphBlock:
    """``runner = CapturedGraphBlock(block, example, image_size, out=fusion[:, 32:])``;
    ``runner(x)`` copies ``x`` into the static input (skipped when ``x`` is None or already the
    static buffer) and replays.  Outputs are static tensors overwritten by every replay.

    ``example`` is a per-pixel feature map ``(B,C,H,W)`` (``kind='feature_map'``) or node features
    ``(B,N,in)`` (``kind='node_features'``).  Weight updates are picked up: the prepared-weight blob
    is refreshed in place before the replay when a parameter changed.

    ``shards`` > 1 records the batch as that many independent sub-batches on parallel branches of the SAME graph (fork /
    join on side streams inside the capture).  Images are independent, so the latency-bound cluster kernel of one
    shard overlaps the HBM-bound pool / un-pool of the other: 192 -> 174 us per step at cfg 2 with two shards (four
    shards lose again to per-kernel overheads).  All shards write into shared full-batch output tensors."""

    def __init__(self, block: GraphBlock, example: torch.Tensor, image_size: Optional[Tuple[int, int]] = None,
                 out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None, want_dense: bool = True,
                 warmup: int = 2, shards: int = 1, packed_small: Optional[torch.Tensor] = None, epilogue=None,
                 epilogue_parallel: bool = False, peer=None):
        """``packed_small``: flat fp32 buffer of ``B*(1 + K*D + N)`` elements; the small per-image outputs
        (``l_partition | region_features | hard_labels``) are then VIEWS of it, i.e. the block kernel writes the
        multi-GPU exchange payload in place.  ``peer``: ``distributed.PeerExchange.slot(i)`` — the block kernel also
        pushes that payload to every rank (needs ``packed_small`` and ``shards == 1``).  ``epilogue(outputs)``: recorded
        at the end of the graph (e.g. ``PeerExchange.wait``), so a step stays ONE driver call.  ``epilogue_parallel``:
        record the epilogue on its own branch right after the block kernel, beside the un-pool (it only needs the small
        outputs)."""
        if not example.is_cuda:
            raise RuntimeError("mingraph_unet_b200 runs on CUDA tensors only (there is no CPU fallback)")
        if block.training and torch.is_grad_enabled():
            raise RuntimeError("CapturedGraphBlock replays the inference path: call block.eval() first")
        self.block = block
        self.kind = "feature_map" if example.dim() == 4 else "node_features"
        self.static_in = example.clone()
        self._kw = dict(image_size=image_size, out=out, out_dtype=out_dtype, want_dense=want_dense)
        dev = example.device
        B = example.shape[0]
        self.shards = max(1, min(int(shards), B))
        self._packed, self._epilogue, self._epi_stream, self._peer = packed_small, epilogue, None, peer
        if peer is not None and (packed_small is None or self.shards != 1):
            raise ValueError("peer needs packed_small and shards == 1 (the push covers the slot's whole batch)")
        self._epi_parallel = bool(epilogue_parallel)
        if self.shards > 1 and not self._shardable(example, image_size, want_dense, out):
            self.shards = 1
        if packed_small is not None and not self._shardable(example, image_size, want_dense, out):
            raise RuntimeError("packed_small needs the one-launch block kernel (shape not supported by mg_block_forward)")
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        if self.shards == 1 and packed_small is None:
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(max(warmup, 1)):
                    self._warm()
            torch.cuda.current_stream(dev).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(self.graph):
                self.outputs: GraphBlockOutput = self._forward()
                if epilogue is not None:
                    epilogue(self.outputs)
            return
        self._alloc_shared_outputs(example, image_size, out, out_dtype, want_dense)
        self._branches = [torch.cuda.Stream(device=dev) for _ in range(self.shards)]
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):
                self._forward_sharded(side)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self._forward_sharded(torch.cuda.current_stream(dev))

    # -- single branch ------------------------------------------------------------------------------
    def _forward(self) -> GraphBlockOutput:
        return self.block(**{self.kind: self.static_in}, **self._kw)

    def _warm(self) -> None:
        out = self._forward()
        if self._epilogue is not None:
            self._epilogue(out)

    # -- parallel shards ----------------------------------------------------------------------------
    def _shardable(self, example, image_size, want_dense, out) -> bool:
        """Shards need the one-launch block kernel (its outputs can be directed into shared buffers)."""
        from . import ops
        blk = self.block
        if not blk.fused:
            return False
        if self.kind == "feature_map":
            H, W = image_size if image_size is not None else tuple(example.shape[-2:])
            in_dim = example.shape[1]
        else:
            if image_size is None:
                return False
            H, W = image_size
            in_dim = example.shape[-1]
        nph, npw = blk.patch_graph_constructor.grid_dims(H, W)
        layers = (blk.patch_gat_model.gat_layers[0], blk.segment_predictor.gnn_predictor.gat_layers[0],
                  blk.region_gat_model.gat_layers[0])
        return bool(ops.block_supported(example.shape[0] // self.shards, nph, npw, in_dim, blk.gat_output_dim,
                                        layers[0].num_heads, layers[1].num_heads, layers[2].num_heads, blk.num_segments))

    def _alloc_shared_outputs(self, example, image_size, out, out_dtype, want_dense) -> None:
        blk, dev = self.block, example.device
        B = example.shape[0]
        H, W = image_size if image_size is not None else tuple(example.shape[-2:])
        nph, npw = blk.patch_graph_constructor.grid_dims(H, W)
        N, K, D = nph * npw, blk.num_segments, blk.gat_output_dim
        f32 = dict(dtype=torch.float32, device=dev)
        self._h = torch.empty((B, N, D), **f32)
        self._S = torch.empty((B, N, K), **f32)
        if self._packed is not None:
            pk = self._packed
            if pk.dtype != torch.float32 or pk.numel() != B * (1 + K * D + N) or not pk.is_contiguous() or pk.device != dev:
                raise ValueError(f"packed_small must be a contiguous float32 buffer of {B * (1 + K * D + N)} elements on {dev}")
            self._loss = pk[:B]
            self._rout = pk[B:B + B * K * D].view(B, K, D)
            self._labels = pk[B + B * K * D:].view(torch.int32).view(B, N)
        else:
            self._labels = torch.empty((B, N), dtype=torch.int32, device=dev)
            self._loss = torch.empty(B, **f32)
            self._rout = torch.empty((B, K, D), **f32)
        dense_dtype = out_dtype if out_dtype is not None else example.dtype
        self._dense = out if (out is not None or not want_dense) else torch.empty((B, D, H, W), dtype=dense_dtype, device=dev)
        self.outputs = GraphBlockOutput(self._dense if want_dense else None, self._loss, self._S, self._labels, self._h,
                                        self._rout, (nph, npw))
        base, rem = divmod(B, self.shards)
        self._ranges, lo = [], 0
        for i in range(self.shards):
            hi = lo + base + (1 if i < rem else 0)
            self._ranges.append((lo, hi))
            lo = hi

    def _forward_sharded(self, main: torch.cuda.Stream) -> None:
        fork = torch.cuda.Event()
        fork.record(main)
        joins, blk_done = [], []

        def after_block():
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(main.device))
            blk_done.append(ev)

        for (lo, hi), st in zip(self._ranges, self._branches):
            st.wait_event(fork)
            with torch.cuda.stream(st):
                kw = dict(self._kw)
                kw["out"] = self._dense[lo:hi] if self._dense is not None else None
                self.block(**{self.kind: self.static_in[lo:hi]}, **kw,
                           _block_outs=(self._h[lo:hi], self._S[lo:hi], self._labels[lo:hi], self._loss[lo:hi], self._rout[lo:hi]),
                           _after_block=after_block if (self._epilogue is not None and self._epi_parallel) else None,
                           _peer=self._peer)
                ev = torch.cuda.Event()
                ev.record(st)
                joins.append(ev)
        if self._epilogue is not None and self._epi_parallel:
            # the epilogue (e.g. the all-gather of the packed small outputs) only needs the block kernels: it runs on its
            # own branch of the graph, in parallel with the HBM-bound un-pool of the same step
            if self._epi_stream is None:
                self._epi_stream = torch.cuda.Stream(device=main.device)
            es = self._epi_stream
            for ev in blk_done:
                es.wait_event(ev)
            with torch.cuda.stream(es):
                self._epilogue(self.outputs)
                ev = torch.cuda.Event()
                ev.record(es)
                joins.append(ev)
        for ev in joins:
            main.wait_event(ev)
        if self._epilogue is not None and not self._epi_parallel:
            with torch.cuda.stream(main):
                self._epilogue(self.outputs)

    def __call__(self, x: Optional[torch.Tensor] = None) -> GraphBlockOutput:
        if x is not None and x.data_ptr() != self.static_in.data_ptr():
            self.static_in.copy_(x, non_blocking=True)
        if self.block.fused:
            self.block._prepared()            # refresh the weight blob in place if a parameter changed
        self.graph.replay()
        return self.outputs


class PipelinedGraphBlock:
    """``depth`` :class:`CapturedGraphBlock` slots (own static input, own outputs, own stream each), used round-robin so
    that CONSECUTIVE steps overlap: while step ``i`` streams its dense map out (HBM-bound un-pool), step ``i+1`` already
    runs its pool and its latency-bound cluster kernel.  Within a step the stage order is a dependency chain
    (pool -> block -> un-pool), so a single replay leaves the HBM pipe idle during the block kernel and the SMs idle
    during the tail of the un-pool; across independent batches nothing has to wait.  Throughput per step approaches the
    step's HBM time; the latency of one step is unchanged.

    ``submit(x)`` enqueues one step on the next slot's stream (after everything the caller has enqueued on the current
    stream so far, so an ``x`` produced there is visible) and returns ``(slot, outputs)``; the outputs are that slot's
    static tensors, valid on ``stream(slot)`` and overwritten ``depth`` submits later.  Consumers either enqueue on
    ``stream(slot)`` or call ``join()`` (current stream waits for every outstanding slot) / ``host_wait(slot)``.
    Weight updates are picked up by the next submit; call ``join()`` before changing parameters."""

    def __init__(self, block: GraphBlock, example: torch.Tensor, image_size: Optional[Tuple[int, int]] = None,
                 outs=None, out_dtype: Optional[torch.dtype] = None, want_dense: bool = True, depth: int = 2,
                 shards: int = 1, warmup: int = 2, packed_small=None, epilogues=None, epilogue_parallel: bool = False,
                 peers=None):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        if outs is not None and len(outs) != depth:
            raise ValueError("outs must hold one dense output (slice) per slot")
        dev = example.device
        self.depth = depth
        self.runners = [CapturedGraphBlock(block, example, image_size, out=None if outs is None else outs[i],
                                           out_dtype=out_dtype, want_dense=want_dense, warmup=warmup, shards=shards,
                                           packed_small=None if packed_small is None else packed_small[i],
                                           epilogue=None if epilogues is None else epilogues[i],
                                           epilogue_parallel=epilogue_parallel,
                                           peer=None if peers is None else peers[i])
                        for i in range(depth)]
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(depth)]
        self.done = [None] * depth
        # events are re-recorded every step instead of created (host time per step matters at ~125 us per step)
        self._ev_in = [torch.cuda.Event() for _ in range(depth)]
        self._ev_done = [torch.cuda.Event() for _ in range(depth)]
        self.shards = self.runners[0].shards
        self._turn = 0

    def stream(self, slot: int) -> torch.cuda.Stream:
        return self.streams[slot]

    @property
    def next_slot(self) -> int:
        """The slot the next ``submit`` will use."""
        return self._turn

    def submit(self, x: Optional[torch.Tensor] = None):
        i = self._turn
        self._turn = (i + 1) % self.depth
        st = self.streams[i]
        ev = self._ev_in[i]
        ev.record(torch.cuda.current_stream(st.device))
        st.wait_event(ev)
        with torch.cuda.stream(st):
            out = self.runners[i](x)
        self.mark(i)
        return i, out

    def mark(self, slot: int) -> None:
        """(Re-)record the slot's completion after the caller enqueued consumer work on ``stream(slot)``."""
        ev = self._ev_done[slot]
        ev.record(self.streams[slot])
        self.done[slot] = ev

    def join(self) -> None:
        cur = torch.cuda.current_stream(self.streams[0].device)
        for ev in self.done:
            if ev is not None:
                cur.wait_event(ev)

    def host_wait(self, slot: int) -> None:
        if self.done[slot] is not None:
            self.done[slot].synchronize()


class CapturedTrainStep:
    """One training step of a :class:`GraphBlock` recorded into CUDA graphs and replayed with one or two driver calls.

    The eager step is host-launch bound (~60 kernels of a few microseconds each plus autograd/optimizer glue).  Recorded:

    * graph A: advance the device-side dropout counter -> patch-mean pool of the static feature map -> block forward
      (training branch, differentiable ops of ``autograd.py``) -> ``loss_fn(out)`` -> ``backward()`` accumulating into ONE
      flat fp32 gradient buffer (every ``p.grad`` is a view of it);
    * between the graphs, with more than one rank: ONE NCCL all-reduce of the flat buffer (22 792 floats for the default
      block), issued eagerly on the same stream;
    * graph B: ``optimizer.step()`` (the optimizer must be built with ``capturable=True``).

    ``dense_cotangent``: the gradient the downstream heads send back into the dense map ``F_g`` (a static ``(B,D,H,W)``
    tensor the caller refreshes in place).  The backward is then seeded with it directly —
    ``autograd.backward([out.f_g, loss_fn(out)], [dense_cotangent, 1])`` — instead of through a scalar formed from the
    dense map, so the block's backward starts at the un-pool's backward scatter with no extra pass over the 134 MB map.

    Gradient plumbing: every ``p.grad`` is dropped before the backward, so autograd hands over the (view of the stacked)
    gradient it computed instead of launching one accumulate kernel per parameter; ONE multi-tensor copy then gathers
    the 20 gradients into the flat buffer (the all-reduce payload), and ``p.grad`` is pointed at the flat views for the
    optimizer.

    With a single rank (or ``allreduce=False``: this rank trains alone) A and B are recorded as one graph.  ``loss_fn`` maps a :class:`GraphBlockOutput` to a scalar and may
    close over other static tensors.  Attention-dropout masks differ from replay to replay (device-side seed addend) and
    are regenerated exactly by the backward of the same replay."""

    def __init__(self, block: GraphBlock, optimizer: torch.optim.Optimizer, example_feature_map: torch.Tensor,
                 image_size: Tuple[int, int], loss_fn, out_dtype: Optional[torch.dtype] = None, group=None, warmup: int = 3,
                 allreduce: bool = True, dense_cotangent: Optional[torch.Tensor] = None):
        import torch.distributed as dist
        from . import ops
        from .autograd import advance_dropout_counter
        if not example_feature_map.is_cuda:
            raise RuntimeError("mingraph_unet_b200 runs on CUDA tensors only (there is no CPU fallback)")
        self.block, self.opt, self.group = block, optimizer, group
        self.world = dist.get_world_size(group) if (allreduce and dist.is_available() and dist.is_initialized()) else 1
        dev = example_feature_map.device
        self.static_in = example_feature_map.clone()
        params = [p for p in block.parameters() if p.requires_grad]
        self.flat_grad = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
        flat_views, off = [], 0
        for p in params:
            flat_views.append(self.flat_grad[off:off + p.numel()].view_as(p))
            off += p.numel()
        ph = block.patch_size

        def fwd_bwd():
            advance_dropout_counter(dev)
            for p in params:
                p.grad = None                      # autograd then hands its gradient over instead of accumulating
            x = ops.pool_patches(self.static_in, ph, ph)
            out = block(node_features=x, image_size=image_size, out_dtype=out_dtype)
            loss = loss_fn(out)
            if dense_cotangent is not None:
                torch.autograd.backward([out.f_g, loss], [dense_cotangent, torch.ones_like(loss)])
            else:
                loss.backward()
            have = [i for i, p in enumerate(params) if p.grad is not None]
            if len(have) != len(params):
                self.flat_grad.zero_()             # (static decision: the same parameters get gradients in every replay)
            torch._foreach_copy_([flat_views[i] for i in have], [params[i].grad for i in have])
            for p, v in zip(params, flat_views):
                p.grad = v
            return loss.detach()

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 3)):          # also fills the graph caches (CSR, backward slot maps)
                fwd_bwd()
                self._allreduce()
                optimizer.step()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph_a = torch.cuda.CUDAGraph()
        self.graph_b = None
        if self.world == 1:
            with torch.cuda.graph(self.graph_a):
                self.loss = fwd_bwd()
                optimizer.step()
        else:
            with torch.cuda.graph(self.graph_a):
                self.loss = fwd_bwd()
            self.graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_b, pool=self.graph_a.pool()):
                optimizer.step()

    def _allreduce(self) -> None:
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
            self.flat_grad.div_(self.world)

    def __call__(self, feature_map: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Copies ``feature_map`` into the static input (skipped when None), runs one step, returns the static loss tensor."""
        if feature_map is not None and feature_map.data_ptr() != self.static_in.data_ptr():
            self.static_in.copy_(feature_map, non_blocking=True)
        self.graph_a.replay()
        if self.graph_b is not None:
            self._allreduce()
            self.graph_b.replay()
        return self.loss
