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
        if att_dropout > 0.0:
            raise NotImplementedError("attention dropout in training mode is not implemented yet; "
                                      "use eval() or dropout_rate=0")
        out, den, z = ops.gat_forward(x, g.rowptr_in, g.col_in, W, a, concat=concat, slope=slope,
                                      nodes_per_graph=g.nodes_per_graph, save=True)
        ctx.save_for_backward(x, W, a, den, z)
        ctx.g, ctx.concat, ctx.slope = g, concat, slope
        return out

    @staticmethod
    def backward(ctx, grad_out):
        raise NotImplementedError("GAT backward kernels are not built yet")


def gat_layer_apply(x: torch.Tensor, g: Graph, W: torch.Tensor, a: torch.Tensor, concat: bool, slope: float,
                    att_dropout: float = 0.0) -> torch.Tensor:
    """Multi-head GAT layer on graph ``g``; differentiable w.r.t. ``x``, ``W``, ``a``."""
    if _wants_grad(x, W, a):
        return _GATLayerFn.apply(x, W, a, g, concat, slope, att_dropout)
    if att_dropout > 0.0:
        raise NotImplementedError("attention dropout in training mode is not implemented yet; "
                                  "use eval() or dropout_rate=0")
    return ops.gat_forward(x, g.rowptr_in, g.col_in, W.detach(), a.detach(), concat=concat, slope=slope,
                           nodes_per_graph=g.nodes_per_graph)


def softmax_rows(logits: torch.Tensor) -> torch.Tensor:
    """``softmax(logits, dim=1)`` (mincut_refinement.py:193)."""
    if _wants_grad(logits):
        return torch.softmax(logits, dim=1)           # stock autograd op until the fused backward lands
    S, _ = ops.softmax_argmax(logits)
    return S


def ncut_loss_apply(h: torch.Tensor, S: torch.Tensor, g: Graph) -> torch.Tensor:
    """Per-graph soft N-cut loss ``(G,)``."""
    if _wants_grad(h, S):
        raise NotImplementedError("N-cut backward kernels are not built yet")
    g.need_out_csr()
    return ops.ncut_loss(h, S, g.rowptr_out, g.col_out, g.nodes_per_graph)
