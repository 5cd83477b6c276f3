"""``torch.autograd.Function`` wrappers: Python owns allocation and saved tensors, the C ABI does
the arithmetic (forward: csrc/gat_forward.cu, ncut.cu; backward: csrc/gat_backward.cu)."""
from __future__ import annotations

import torch

from . import ops
from .graph import Graph


def _wants_grad(*ts) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts)


class _GATLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, a, g: Graph, concat: bool, slope: float, att_dropout: float):
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if att_dropout > 0.0 else 0     # CPU generator: no device sync
        out, den, z = ops.gat_forward(x, g.rowptr_in, g.col_in, W, a, concat=concat, slope=slope,
                                      nodes_per_graph=g.nodes_per_graph, save=True, dropout_p=att_dropout, seed=seed)
        ctx.save_for_backward(x, W, a, den, z)
        ctx.g, ctx.concat, ctx.slope, ctx.p, ctx.seed = g, concat, slope, att_dropout, seed
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, W, a, den, z = ctx.saved_tensors
        g = ctx.g
        g.need_backward_maps()
        gx, gW, ga = ops.gat_backward(x, g.rowptr_in, g.col_in, g.rowptr_out, g.col_out, g.slot_out2in, W, a, den, z,
                                      grad_out, concat=ctx.concat, slope=ctx.slope, nodes_per_graph=g.nodes_per_graph,
                                      dropout_p=ctx.p, seed=ctx.seed)
        return gx.to(x.dtype), gW.to(W.dtype), ga.to(a.dtype), None, None, None, None


def gat_layer_apply(x: torch.Tensor, g: Graph, W: torch.Tensor, a: torch.Tensor, concat: bool, slope: float,
                    att_dropout: float = 0.0) -> torch.Tensor:
    """Multi-head GAT layer on graph ``g``; differentiable w.r.t. ``x``, ``W``, ``a``."""
    if _wants_grad(x, W, a):
        return _GATLayerFn.apply(x, W, a, g, concat, slope, att_dropout)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if att_dropout > 0.0 else 0
    return ops.gat_forward(x, g.rowptr_in, g.col_in, W.detach(), a.detach(), concat=concat, slope=slope,
                           nodes_per_graph=g.nodes_per_graph, dropout_p=att_dropout, seed=seed)


def softmax_rows(logits: torch.Tensor) -> torch.Tensor:
    """``softmax(logits, dim=1)`` (mincut_refinement.py:193)."""
    if _wants_grad(logits):
        return torch.softmax(logits, dim=1)           # stock autograd op until the fused backward lands
    S, _ = ops.softmax_argmax(logits)
    return S


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


def ncut_loss_apply(h: torch.Tensor, S: torch.Tensor, g: Graph) -> torch.Tensor:
    """Per-graph soft N-cut loss ``(G,)``; differentiable w.r.t. ``h`` and ``S``."""
    g.need_out_csr()
    if _wants_grad(h, S):
        return _NcutFn.apply(h.contiguous(), S.contiguous(), g)
    return ops.ncut_loss(h, S, g.rowptr_out, g.col_out, g.nodes_per_graph)
