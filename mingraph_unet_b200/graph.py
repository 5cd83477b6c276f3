"""Device-side graph handles: COO ``edge_index`` (the reference's interchange format) plus the
int32 CSR views the kernels consume.

The reference passes ``edge_index (2,E) int64`` between its modules
(preprocessing/graph_construction/patch_graph_construction.py:60-61,97).  The kernels want a CSR
sorted by target (forward gather) and by source (N-cut degree, backward).  ``Graph`` objects are
cached per ``edge_index`` tensor so a graph is sorted once, and the grid / complete graphs built
by this package carry their closed-form CSR from the start (no sort at all).
"""
from __future__ import annotations

import weakref
from typing import Dict, Optional, Tuple

import torch

from . import ops


class Graph:
    """A (possibly block-diagonal) directed graph on ``N`` nodes.

    in-CSR: ``rowptr_in[j]..rowptr_in[j+1]`` lists the sources of edges into ``j``;
    out-CSR: ``rowptr_out[i]..`` lists the targets of edges out of ``i``; both keep ascending COO
    edge id inside a row.  ``nodes_per_graph`` > 0 marks a batch of equal-size independent graphs.
    """

    def __init__(self, N: int, E: int, device, nodes_per_graph: int = 0):
        self.N, self.E, self.device, self.nodes_per_graph = N, E, device, nodes_per_graph
        self.edge_index: Optional[torch.Tensor] = None
        self.rowptr_in = self.col_in = self.eid_in = None
        self.rowptr_out = self.col_out = self.eid_out = None
        self.symmetric_csr = False
        self.slot_out2in = None

    # -- constructors ---------------------------------------------------------------------------
    @classmethod
    def grid(cls, Hp: int, Wp: int, device, B: int = 1, with_edge_index: bool = False) -> "Graph":
        key = ("grid", Hp, Wp, B, str(device))
        g = _STATIC.get(key)
        if g is None:
            N, E = Hp * Wp, ops.grid_num_edges(Hp, Wp)
            g = cls(B * N, B * E, device, nodes_per_graph=N if B > 1 else 0)
            g.rowptr_in, g.col_in, g.eid_in, g.eid_out = ops.grid_csr(Hp, Wp, device, B, with_eid=True)
            if B > 1:      # eids are per-image COO ids: offset them so they index the batched edge list
                off = (torch.arange(B, device=device, dtype=torch.int32) * E).repeat_interleave(E)
                g.eid_in, g.eid_out = g.eid_in + off, g.eid_out + off
            g.rowptr_out, g.col_out = g.rowptr_in, g.col_in       # the grid is symmetric
            g.symmetric_csr = True
            g.per_graph_edges = E
            _STATIC[key] = g
        if with_edge_index and g.edge_index is None:
            g.edge_index = ops.grid_edge_index(Hp, Wp, device, B, offset_nodes=B > 1)
        return g

    @classmethod
    def complete(cls, K: int, device, B: int = 1) -> "Graph":
        key = ("complete", K, B, str(device))
        g = _STATIC.get(key)
        if g is None:
            g = cls(B * K, B * K * (K - 1), device, nodes_per_graph=K if B > 1 else 0)
            g.edge_index = ops.complete_edge_index(K, device, B, offset_nodes=True)
            g.rowptr_in, g.col_in = ops.complete_csr(K, device, B)
            g.rowptr_out, g.col_out = g.rowptr_in, g.col_in
            g.symmetric_csr = True
            g.per_graph_edges = K * (K - 1)
            _STATIC[key] = g
        return g

    @classmethod
    def from_edge_index(cls, edge_index: torch.Tensor, N: int) -> "Graph":
        """Cached per tensor object (and its in-place version counter)."""
        key = id(edge_index)
        hit = _BY_TENSOR.get(key)
        if hit is not None:
            ref, ver, n, g = hit
            if ref() is edge_index and ver == edge_index._version and n == N:
                return g
        if edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise ValueError("edge_index must have shape (2, E)")
        g = cls(N, int(edge_index.shape[1]), edge_index.device)
        g.edge_index = edge_index
        if g.E > 0:
            # a caller-supplied edge list: node ids outside [0, N) raise IndexError like the reference's indexing does
            # (one device sync per NEW edge_index tensor, amortised by this cache; not possible while a CUDA graph is
            # being captured, where the caller is responsible for having validated the tensor before)
            check = edge_index.is_cuda and not torch.cuda.is_current_stream_capturing()
            g.rowptr_in, g.col_in, g.eid_in = ops.csr_from_coo(edge_index, N, by_target=True, check=check)
        register(edge_index, g)
        return g

    def need_out_csr(self) -> None:
        if self.rowptr_out is None:
            self.rowptr_out, self.col_out, self.eid_out = ops.csr_from_coo(self.edge_index, self.N, by_target=False)

    def need_backward_maps(self) -> None:
        """Both CSR views plus, per out-CSR slot, the in-CSR slot of the same edge (GAT backward)."""
        if self.slot_out2in is not None:
            return
        if self.eid_in is None or self.eid_out is None:
            if self.edge_index is None:
                raise RuntimeError("graph has no edge_index to derive the backward maps from")
            self.rowptr_in, self.col_in, self.eid_in = ops.csr_from_coo(self.edge_index, self.N, by_target=True)
            self.rowptr_out, self.col_out, self.eid_out = ops.csr_from_coo(self.edge_index, self.N, by_target=False)
        self.slot_out2in = ops.edge_slot_map(self.eid_in, self.eid_out)


_STATIC: Dict[tuple, Graph] = {}
_BY_TENSOR: Dict[int, Tuple[weakref.ref, int, int, Graph]] = {}


def register(edge_index: torch.Tensor, g: Graph) -> None:
    """Attach a ready-made Graph (e.g. closed-form grid CSR) to an ``edge_index`` tensor."""
    key = id(edge_index)

    def _drop(_ref, key=key):
        _BY_TENSOR.pop(key, None)

    _BY_TENSOR[key] = (weakref.ref(edge_index, _drop), edge_index._version, g.N, g)


def clear_caches() -> None:
    _STATIC.clear()
    _BY_TENSOR.clear()
