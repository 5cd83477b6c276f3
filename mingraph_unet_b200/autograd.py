"""``torch.autograd.Function`` wrappers: Python owns allocation and saved tensors, the C ABI does
the arithmetic (forward: csrc/gat_forward.cu, ncut.cu; backward: csrc/gat_backward.cu)."""
from __future__ import annotations

import torch

from . import ops
from .graph import Graph


_DROPOUT_COUNTERS = {}


def dropout_counter(device) -> torch.Tensor:
    """Persistent device-side counter added to every attention-dropout seed when the kernels run.  A captured training
    step advances it inside the graph (:func:`advance_dropout_counter`), so each replay draws fresh masks while the
    backward of the same replay regenerates exactly the forward's."""
    key = str(device)
    t = _DROPOUT_COUNTERS.get(key)
    if t is None:
        t = torch.zeros(1, dtype=torch.int64, device=device)
        _DROPOUT_COUNTERS[key] = t
    return t


def advance_dropout_counter(device) -> None:
    dropout_counter(device).add_(1)


def _wants_grad(*ts) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts)


class _StackHeadsFn(torch.autograd.Function):
    """``stack([p.reshape(shape) for p in params])`` of the per-head parameters (the reference keeps one ``nn.Linear``
    per head, graph_attention.py:28-31,144-148; the kernels take all heads stacked).  The backward hands every head a
    VIEW of the stacked gradient: no unbind / reshape / accumulate kernels (torch's own ``stack`` + ``view`` graph cost
    ~60 tiny launches per training step for the block's 20 parameters)."""

    @staticmethod
    def forward(ctx, shape, *params):
        ctx.shapes = [p.shape for p in params]
        return torch.stack([p.reshape(shape) for p in params], 0)

    @staticmethod
    def backward(ctx, grad):
        return (None,) + tuple(grad[i].view(s) for i, s in enumerate(ctx.shapes))


def stack_heads(params, shape) -> torch.Tensor:
    """``(H, *shape)`` stack of per-head parameters; differentiable (gradients arrive as views, see above)."""
    params = list(params)
    if _wants_grad(*params):
        return _StackHeadsFn.apply(tuple(shape), *params)
    return torch.stack([p.detach().reshape(shape) for p in params], 0)


class _GATLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, a, g: Graph, concat: bool, slope: float, att_dropout: float, out_dtype=None):
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if att_dropout > 0.0 else 0     # CPU generator: no device sync
        seed_dev = dropout_counter(x.device) if att_dropout > 0.0 else None
        out, den, z, work = ops.gat_forward(x, g.rowptr_in, g.col_in, W, a, concat=concat, slope=slope,
                                            nodes_per_graph=g.nodes_per_graph, save=True, dropout_p=att_dropout, seed=seed,
                                            out_dtype=out_dtype, seed_dev=seed_dev)
        ctx.save_for_backward(x, W, a, den, z, work)       # work: the forward's attention scalars, reused by the backward
        ctx.g, ctx.concat, ctx.slope, ctx.p, ctx.seed, ctx.seed_dev = g, concat, slope, att_dropout, seed, seed_dev
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, W, a, den, z, work = ctx.saved_tensors
        g = ctx.g
        g.need_backward_maps()
        gx, gW, ga = ops.gat_backward(x, g.rowptr_in, g.col_in, g.rowptr_out, g.col_out, g.slot_out2in, W, a, den, z,
                                      grad_out, concat=ctx.concat, slope=ctx.slope, nodes_per_graph=g.nodes_per_graph,
                                      dropout_p=ctx.p, seed=ctx.seed, seed_dev=ctx.seed_dev, fwd_work=work)
        return gx.to(x.dtype), gW.to(W.dtype), ga.to(a.dtype), None, None, None, None, None


def gat_layer_apply(x: torch.Tensor, g: Graph, W: torch.Tensor, a: torch.Tensor, concat: bool, slope: float,
                    att_dropout: float = 0.0, out_dtype=None) -> torch.Tensor:
    """Multi-head GAT layer on graph ``g``; differentiable w.r.t. ``x``, ``W``, ``a``."""
    if _wants_grad(x, W, a):
        return _GATLayerFn.apply(x, W, a, g, concat, slope, att_dropout, out_dtype)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if att_dropout > 0.0 else 0
    return ops.gat_forward(x, g.rowptr_in, g.col_in, W.detach(), a.detach(), concat=concat, slope=slope,
                           nodes_per_graph=g.nodes_per_graph, dropout_p=att_dropout, seed=seed, out_dtype=out_dtype,
                           seed_dev=dropout_counter(x.device) if att_dropout > 0.0 else None)


class _SoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits):
        S, labels = ops.softmax_argmax(logits)
        ctx.save_for_backward(S)
        ctx.mark_non_differentiable(labels)
        return S, labels

    @staticmethod
    def backward(ctx, grad_S, _grad_labels):
        (S,) = ctx.saved_tensors
        return ops.softmax_backward(S, grad_S)


def softmax_rows_with_labels(logits: torch.Tensor):
    """``(softmax(logits, 1), argmax)`` in one kernel; differentiable in the first output."""
    if _wants_grad(logits):
        return _SoftmaxFn.apply(logits.contiguous())
    return ops.softmax_argmax(logits)


def softmax_rows(logits: torch.Tensor) -> torch.Tensor:
    """``softmax(logits, dim=1)`` (mincut_refinement.py:193)."""
    if _wants_grad(logits):
        return _SoftmaxFn.apply(logits.contiguous())[0]
    S, _ = ops.softmax_argmax(logits)
    return S


class _SegmentMeanFn(torch.autograd.Function):
    """Region mean pool (train_end_to_end.py:368-373); labels are data (argmax), not differentiated."""

    @staticmethod
    def forward(ctx, h, labels, K: int):
        out, counts = ops.segment_mean(h, labels, K, with_counts=True)
        ctx.save_for_backward(labels, counts)
        ctx.N = h.shape[1]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        labels, counts = ctx.saved_tensors
        return ops.segment_mean_backward(grad_out, labels, counts, ctx.N), None, None


def segment_mean_apply(h: torch.Tensor, labels: torch.Tensor, K: int) -> torch.Tensor:
    if _wants_grad(h):
        return _SegmentMeanFn.apply(h.contiguous(), labels, K)
    return ops.segment_mean(h, labels, K)


class _UnpoolFn(torch.autograd.Function):
    """Gather by label + nearest up-sampling (train_end_to_end.py:403-421); the backward is the segmented
    reduction of the dense gradient back onto the K region rows."""

    @staticmethod
    def forward(ctx, out, table, labels, Hp, Wp, H, W, out_dtype):
        # `out` (the tensor modified in place) is the FIRST argument: autograd's CopySlices, which wraps this function when
        # `out` is a view of a larger buffer, takes input 0 to be the modified tensor
        res = ops.unpool_nearest(table, labels, Hp, Wp, H, W, out=out, out_dtype=out_dtype)
        ctx.save_for_backward(labels)
        ctx.dims = (table.shape[1], Hp, Wp)
        if out is not None:
            ctx.mark_dirty(out)
        return res

    @staticmethod
    def backward(ctx, grad_out):
        (labels,) = ctx.saved_tensors
        K, Hp, Wp = ctx.dims
        # gradient w.r.t. the previous content of `out`: none (it was overwritten; CopySlices reads None as zeros)
        return None, ops.unpool_nearest_backward(grad_out, labels, K, Hp, Wp), None, None, None, None, None, None


def unpool_apply(table, labels, Hp, Wp, H, W, out=None, out_dtype=torch.float32):
    if _wants_grad(table):
        if out is not None:
            # writing into a caller-owned slice (the fusion buffer) under autograd: an in-place op on `out` (mark_dirty),
            # so a head that reads the BUFFER back-propagates into the region rows — the gradient path is the buffer's
            # own history, not a detached copy
            if out.requires_grad and out.is_leaf:
                raise RuntimeError("unpool_apply: `out` is a leaf that requires grad; pass a buffer (slice) that does not")
            return _UnpoolFn.apply(out, table.contiguous(), labels, Hp, Wp, H, W, out.dtype)
        return _UnpoolFn.apply(None, table.contiguous(), labels, Hp, Wp, H, W, out_dtype)
    return ops.unpool_nearest(table, labels, Hp, Wp, H, W, out=out, out_dtype=out_dtype)


class _NcutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, S, g: Graph):
        loss, stats = ops.ncut_loss(h, S, g.rowptr_out, g.col_out, g.nodes_per_graph, with_stats=True)
        ctx.save_for_backward(h, S, stats)
        ctx.g = g
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        h, S, stats = ctx.saved_tensors
        g = ctx.g
        gh, gS = ops.ncut_backward(h, S, g.rowptr_out, g.col_out, g.rowptr_in, g.col_in, stats, grad_loss,
                                   g.nodes_per_graph)
        return gh, gS, None


class _EdgeWeightFn(torch.autograd.Function):
    """``w_e = exp(-|h_src - h_tgt|^2 / 2)`` (mincut_refinement.py:43-51), differentiable w.r.t. ``h`` like the reference
    method.  The backward of this stand-alone accessor is index glue in torch (the block's own N-cut backward is
    ``mg_ncut_backward``)."""

    @staticmethod
    def forward(ctx, h, edge_index):
        w = ops.ncut_edge_weights(h, edge_index)
        ctx.save_for_backward(h, edge_index, w)
        return w

    @staticmethod
    def backward(ctx, grad_w):
        h, ei, w = ctx.saved_tensors
        c = (grad_w * w).unsqueeze(1) * (h[ei[0]] - h[ei[1]])          # d w_e / d h_src = -w_e (h_src - h_tgt)
        gh = torch.zeros_like(h)
        gh.index_add_(0, ei[0], -c)
        gh.index_add_(0, ei[1], c)
        return gh, None


def edge_weights_apply(h: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    if _wants_grad(h):
        return _EdgeWeightFn.apply(h.contiguous(), edge_index)
    return ops.ncut_edge_weights(h, edge_index)


def ncut_loss_apply(h: torch.Tensor, S: torch.Tensor, g: Graph) -> torch.Tensor:
    """Per-graph soft N-cut loss ``(G,)``; differentiable w.r.t. ``h`` and ``S``."""
    g.need_out_csr()
    if _wants_grad(h, S):
        D = h.shape[-1]
        if D % 4 != 0 or D > 128:       # what mg_ncut_backward takes: fail at the forward, not in the middle of backward()
            raise RuntimeError(f"the N-cut backward kernel needs a feature width that is a multiple of 4 and <= 128 (got {D})")
        return _NcutFn.apply(h.contiguous(), S.contiguous(), g)
    return ops.ncut_loss(h, S, g.rowptr_out, g.col_out, g.nodes_per_graph)


class _FeatureLossFn(torch.autograd.Function):
    """FeatureConsistencyLoss (model/unet/feature_loss.py:103-123); labels are data, not differentiated."""

    @staticmethod
    def forward(ctx, f_unet, f_graph, y, margin: float):
        ctx.save_for_backward(f_unet, f_graph, y)
        ctx.margin = margin
        return ops.feature_consistency_loss(f_unet, f_graph, y, margin)

    @staticmethod
    def backward(ctx, grad_loss):
        f_unet, f_graph, y = ctx.saved_tensors
        gu, gg = ops.feature_consistency_loss_backward(f_unet, f_graph, y, ctx.margin, grad_loss,
                                                       ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return (None if gu is None else gu.to(f_unet.dtype)), (None if gg is None else gg.to(f_graph.dtype)), None, None


def feature_loss_apply(f_unet: torch.Tensor, f_graph: torch.Tensor, y: torch.Tensor, margin: float) -> torch.Tensor:
    if _wants_grad(f_unet, f_graph):
        return _FeatureLossFn.apply(f_unet, f_graph, y, margin)
    return ops.feature_consistency_loss(f_unet, f_graph, y, margin)


class _TVLossFn(torch.autograd.Function):
    """TVLoss (scripts/train_end_to_end.py:73-89)."""

    @staticmethod
    def forward(ctx, x, weight: float):
        ctx.save_for_backward(x)
        ctx.weight = weight
        return ops.tv_loss(x, weight).clone()

    @staticmethod
    def backward(ctx, grad_loss):
        (x,) = ctx.saved_tensors
        return ops.tv_loss_backward(x, ctx.weight, grad_loss).to(x.dtype), None


def tv_loss_apply(x: torch.Tensor, weight: float) -> torch.Tensor:
    if _wants_grad(x):
        return _TVLossFn.apply(x, weight)
    return ops.tv_loss(x, weight)
