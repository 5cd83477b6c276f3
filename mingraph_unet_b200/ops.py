"""Tensor-level wrappers over the C ABI (``include/mingraph_b200.h``).

PyTorch is used here only as the owner of device memory and of the current CUDA stream;
all arithmetic happens in ``libmingraph_b200.so``.  Every function requires CUDA tensors and
raises otherwise (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import MG_BF16, MG_F32, call

_DT = {torch.float32: MG_F32, torch.bfloat16: MG_BF16}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts: torch.Tensor) -> torch.device:
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("mingraph_unet_b200 runs on CUDA tensors only (there is no CPU fallback)")
        if dev is not None and t.device != dev:
            raise RuntimeError("tensors live on different devices")
        dev = t.device
    return dev


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _dtype_code(dt: torch.dtype) -> int:
    try:
        return _DT[dt]
    except KeyError:
        raise RuntimeError(f"unsupported dtype {dt}: the graph block handles float32 and bfloat16") from None


# ---------------------------------------------------------------------------------------------
# graph construction
# ---------------------------------------------------------------------------------------------
def grid_num_edges(Hp: int, Wp: int) -> int:
    return int(_lib.load().mg_grid_num_edges(Hp, Wp))


def grid_edge_index(Hp: int, Wp: int, device, B: int = 1, offset_nodes: bool = False) -> torch.Tensor:
    """``(2, B*E)`` int64 COO of the 4-connected patch grid in the reference's edge order
    (patch_graph_construction.py:78-97)."""
    E = grid_num_edges(Hp, Wp)
    ei = torch.empty((2, B * E), dtype=torch.int64, device=device)
    with torch.cuda.device(ei.device):
        call("mg_grid_edge_index", Hp, Wp, B, int(offset_nodes), ei.data_ptr(), _stream())
    return ei


def grid_csr(Hp: int, Wp: int, device, B: int = 1, with_eid: bool = False):
    """Closed-form block-diagonal CSR of ``B`` grid graphs: ``rowptr (B*N+1)``, ``col (B*E)`` int32."""
    N, E = Hp * Wp, grid_num_edges(Hp, Wp)
    rowptr = torch.empty(B * N + 1, dtype=torch.int32, device=device)
    col = torch.empty(max(B * E, 1), dtype=torch.int32, device=device)
    eid_in = torch.empty_like(col) if with_eid else None
    eid_out = torch.empty_like(col) if with_eid else None
    with torch.cuda.device(rowptr.device):
        call("mg_grid_csr", Hp, Wp, B, rowptr.data_ptr(), col.data_ptr(), _ptr(eid_in), _ptr(eid_out), _stream())
    col = col[: B * E]
    if with_eid:
        return rowptr, col, eid_in[: B * E], eid_out[: B * E]
    return rowptr, col


def complete_edge_index(K: int, device, B: int = 1, offset_nodes: bool = False) -> torch.Tensor:
    """Complete digraph on K regions, ``triu`` pairs then reversed (train_end_to_end.py:376-380)."""
    E = K * (K - 1)
    ei = torch.empty((2, B * E), dtype=torch.int64, device=device)
    if E > 0:
        with torch.cuda.device(ei.device):
            call("mg_complete_edge_index", K, B, int(offset_nodes), ei.data_ptr(), _stream())
    return ei


def complete_csr(K: int, device, B: int = 1):
    rowptr = torch.empty(B * K + 1, dtype=torch.int32, device=device)
    col = torch.empty(max(B * K * (K - 1), 1), dtype=torch.int32, device=device)
    with torch.cuda.device(rowptr.device):
        call("mg_complete_csr", K, B, rowptr.data_ptr(), col.data_ptr(), _stream())
    return rowptr, col[: B * K * (K - 1)]


def csr_from_coo(edge_index: torch.Tensor, N: int, by_target: bool = True, check: bool = False):
    """Stable CSR of a caller-supplied ``(2,E)`` int64 ``edge_index`` (ascending COO id inside each
    row).  Returns ``rowptr, col, eid``.  ``check=True`` synchronises and raises ``IndexError`` if an
    index is outside ``[0,N)`` like the reference's indexing would."""
    _need_cuda(edge_index)
    if edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise ValueError("edge_index must have shape (2, E)")
    if edge_index.dtype != torch.int64:
        edge_index = edge_index.long()
    edge_index = edge_index.contiguous()
    E = edge_index.shape[1]
    dev = edge_index.device
    rowptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    col = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
    eid = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
    work = torch.empty(int(_lib.load().mg_csr_work_bytes(N, E)), dtype=torch.uint8, device=dev)
    status = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        call("mg_csr_from_coo", edge_index.data_ptr(), E, N, int(by_target), rowptr.data_ptr(), col.data_ptr(),
             eid.data_ptr(), work.data_ptr(), status.data_ptr(), _stream())
    if check and int(status.item()) != 0:
        raise IndexError(f"edge_index holds node ids outside [0, {N})")
    return rowptr, col[:E], eid[:E]


def knn_graph(x: torch.Tensor, k: int, nodes_per_graph: int = 0, with_dist: bool = False):
    """k nearest neighbours of every node within its graph (squared Euclidean, fp32, ties to the lower id, self
    excluded).  Returns ``edge_index (2, N*k) int64`` (row 0 = neighbour/source, row 1 = node/target), ``rowptr``,
    ``col`` (in-CSR, int32) and optionally the distances ``(N, k)``.  Not part of the reference (csrc/knn.cu)."""
    _need_cuda(x)
    if x.dim() != 2 or x.dtype != torch.float32:
        raise RuntimeError("knn_graph expects float32 node features (N, D)")
    x = x.contiguous()
    N, D = x.shape
    ei = torch.empty((2, N * k), dtype=torch.int64, device=x.device)
    rowptr = torch.empty(N + 1, dtype=torch.int32, device=x.device)
    col = torch.empty(N * k, dtype=torch.int32, device=x.device)
    dist = torch.empty((N, k), dtype=torch.float32, device=x.device) if with_dist else None
    with torch.cuda.device(x.device):
        call("mg_knn_graph", x.data_ptr(), N, D, int(k), int(nodes_per_graph), ei.data_ptr(), rowptr.data_ptr(), col.data_ptr(),
             _ptr(dist), _stream())
    return (ei, rowptr, col, dist) if with_dist else (ei, rowptr, col)


# ---------------------------------------------------------------------------------------------
# pooling / un-pooling
# ---------------------------------------------------------------------------------------------
def pool_patches(x: torch.Tensor, ph: int, pw: Optional[int] = None, out_dtype: Optional[torch.dtype] = None):
    """Patch mean pool ``(B,C,Hf,Wf) -> (B, Hp*Wp, C)`` (zero padded right/bottom, divisor ph*pw)."""
    _need_cuda(x)
    if x.dim() != 4:
        raise ValueError("pool_patches expects (B, C, Hf, Wf)")
    pw = ph if pw is None else pw
    x = x.contiguous()
    B, Cc, Hf, Wf = x.shape
    Hp, Wp = -(-Hf // ph), -(-Wf // pw)
    out_dtype = x.dtype if out_dtype is None else out_dtype
    out = torch.empty((B, Hp * Wp, Cc), dtype=out_dtype, device=x.device)
    with torch.cuda.device(x.device):
        call("mg_pool_patches", x.data_ptr(), _dtype_code(x.dtype), B, Cc, Hf, Wf, ph, pw, out.data_ptr(),
             _dtype_code(out_dtype), _stream())
    return out


def segment_mean(h: torch.Tensor, labels: torch.Tensor, K: int, with_counts: bool = False):
    """Region mean pool ``(B,N,D),(B,N) int32 -> (B,K,D)``; empty regions give zeros."""
    _need_cuda(h, labels)
    if h.dtype != torch.float32 or labels.dtype != torch.int32:
        raise RuntimeError("segment_mean expects float32 features and int32 labels")
    h, labels = h.contiguous(), labels.contiguous()
    B, N, D = h.shape
    out = torch.empty((B, K, D), dtype=torch.float32, device=h.device)
    counts = torch.empty((B, K), dtype=torch.int32, device=h.device) if with_counts else None
    work = torch.empty(int(_lib.load().mg_segment_work_bytes(B, N, D, K)), dtype=torch.uint8, device=h.device)
    with torch.cuda.device(h.device):
        call("mg_segment_mean", h.data_ptr(), labels.data_ptr(), B, N, D, K, out.data_ptr(), _ptr(counts), work.data_ptr(),
             _stream())
    return (out, counts) if with_counts else out


def unpool_nearest(table: torch.Tensor, labels: Optional[torch.Tensor], Hp: int, Wp: int, H: int, W: int,
                   out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """``out[b,d,y,x] = table[b, labels[b, iy*Wp+ix], d]`` with torch's nearest index rule.
    ``out`` may be a channel slice ``buf[:, c0:c0+D]`` of a contiguous ``(B,Ctot,H,W)`` buffer."""
    _need_cuda(table, labels, out)
    if table.dtype != torch.float32:
        raise RuntimeError("unpool_nearest expects a float32 table")
    table = table.contiguous()
    B, K, D = table.shape
    if labels is not None:
        if labels.dtype != torch.int32:
            raise RuntimeError("labels must be int32")
        labels = labels.contiguous()
    if out is None:
        out = torch.empty((B, D, H, W), dtype=out_dtype, device=table.device)
    if tuple(out.shape) != (B, D, H, W) or out.stride()[1:] != (H * W, W, 1):
        raise ValueError("out must be (B,D,H,W) with contiguous (D,H,W) planes")
    with torch.cuda.device(table.device):
        call("mg_unpool_nearest", table.data_ptr(), _ptr(labels), B, K, D, Hp, Wp, H, W, out.data_ptr(),
             _dtype_code(out.dtype), out.stride(0) if B > 1 else D * H * W, _stream())
    return out


# ---------------------------------------------------------------------------------------------
# graph attention
# ---------------------------------------------------------------------------------------------
def gat_forward(x: torch.Tensor, rowptr: torch.Tensor, col: torch.Tensor, W: torch.Tensor, a: torch.Tensor,
                concat: bool = False, slope: float = 0.2, nodes_per_graph: int = 0,
                out_dtype: Optional[torch.dtype] = None, save: bool = False, dropout_p: float = 0.0, seed: int = 0,
                seed_dev: Optional[torch.Tensor] = None):
    """Multi-head GAT layer forward (eval semantics).  ``x (N,in)`` f32|bf16, ``W (H,F,in)``,
    ``a (H,2F)`` f32, in-CSR ``rowptr/col`` int32.  Returns ``out`` or ``(out, den, z, work)`` if ``save`` (``work``: the
    call's workspace, whose attention scalars ``gat_backward(fwd_work=...)`` reuses)."""
    _need_cuda(x, rowptr, col, W, a)
    if x.dim() != 2 or W.dim() != 3 or a.dim() != 2:
        raise ValueError("gat_forward expects x (N,in), W (H,F,in), a (H,2F)")
    N, in_dim = x.shape
    heads, F, in_w = W.shape
    if in_w != in_dim or tuple(a.shape) != (heads, 2 * F):
        raise ValueError(f"weight shapes {tuple(W.shape)} / {tuple(a.shape)} do not match x {tuple(x.shape)}")
    E = col.numel()
    if E == 0:
        # the reference fails in torch.max(e) on an empty tensor (graph_attention.py:86)
        raise RuntimeError("gat_forward: edge_index is empty (max() of an empty edge set)")
    x = x.contiguous()
    W = W.contiguous().float()
    a = a.contiguous().float()
    out_dtype = x.dtype if out_dtype is None else out_dtype
    out = torch.empty((N, heads * F if concat else F), dtype=out_dtype, device=x.device)
    G = N // nodes_per_graph if nodes_per_graph > 0 else 1
    lib = _lib.load()
    work = torch.empty(max(int(lib.mg_gat_work_bytes(N, in_dim, F, heads, G)), 256), dtype=torch.uint8, device=x.device)
    den = torch.empty((N, heads), dtype=torch.float32, device=x.device) if save else None
    z = torch.empty((N, heads, in_dim), dtype=torch.float32, device=x.device) if save else None
    with torch.cuda.device(x.device):
        call("mg_gat_forward", x.data_ptr(), _dtype_code(x.dtype), rowptr.data_ptr(), col.data_ptr(), N, E,
             W.data_ptr(), a.data_ptr(), in_dim, F, heads, int(concat), float(slope), int(nodes_per_graph),
             float(dropout_p), int(seed), _ptr(seed_dev), out.data_ptr(), _dtype_code(out_dtype), work.data_ptr(), _ptr(den),
             _ptr(z), _stream())
    return (out, den, z, work) if save else out


def edge_slot_map(eid_in: torch.Tensor, eid_out: torch.Tensor) -> torch.Tensor:
    """For every out-CSR slot, the in-CSR slot holding the same edge (needed by ``gat_backward``)."""
    _need_cuda(eid_in, eid_out)
    E = eid_in.numel()
    out = torch.empty(max(E, 1), dtype=torch.int32, device=eid_in.device)
    work = torch.empty(max(E, 1), dtype=torch.int32, device=eid_in.device)
    with torch.cuda.device(eid_in.device):
        call("mg_edge_slot_map", eid_in.data_ptr(), eid_out.data_ptr(), E, work.data_ptr(), out.data_ptr(), _stream())
    return out[:E]


def gat_backward(x, rowptr_in, col_in, rowptr_out, col_out, slot_out2in, W, a, den, z, grad_out, concat: bool = False,
                 slope: float = 0.2, nodes_per_graph: int = 0, dropout_p: float = 0.0, seed: int = 0,
                 seed_dev: Optional[torch.Tensor] = None, fwd_work: Optional[torch.Tensor] = None):
    """Backward of ``gat_forward``: returns ``grad_x (N,in) f32, grad_W (H,F,in), grad_a (H,2F)``.  ``fwd_work``: the
    forward call's workspace (``gat_forward(save=True)``): its scores / maxima are reused (3 launches fewer)."""
    _need_cuda(x, W, a, den, z, grad_out)
    N, in_dim = x.shape
    heads, F, _ = W.shape
    E = col_in.numel()
    x = x.contiguous()
    W = W.contiguous().float()
    a = a.contiguous().float()
    grad_out = grad_out.contiguous().float()
    dev = x.device
    gx = torch.empty((N, in_dim), dtype=torch.float32, device=dev)
    gW = torch.empty((heads, F, in_dim), dtype=torch.float32, device=dev)
    ga = torch.empty((heads, 2 * F), dtype=torch.float32, device=dev)
    G = N // nodes_per_graph if nodes_per_graph > 0 else 1
    work = torch.empty(max(int(_lib.load().mg_gat_backward_work_bytes(N, E, in_dim, F, heads, G)), 256), dtype=torch.uint8,
                       device=dev)
    with torch.cuda.device(dev):
        call("mg_gat_backward", x.data_ptr(), _dtype_code(x.dtype), rowptr_in.data_ptr(), col_in.data_ptr(),
             rowptr_out.data_ptr(), col_out.data_ptr(), slot_out2in.data_ptr(), N, E, W.data_ptr(), a.data_ptr(), in_dim, F,
             heads, int(concat), float(slope), int(nodes_per_graph), float(dropout_p), int(seed), _ptr(seed_dev), den.data_ptr(),
             z.data_ptr(), grad_out.data_ptr(), gx.data_ptr(), gW.data_ptr(), ga.data_ptr(), work.data_ptr(), _ptr(fwd_work),
             _stream())
    return gx, gW, ga


def softmax_argmax(logits: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Row softmax and first-max argmax of ``(N,K)`` float32 logits -> ``S (N,K)``, ``labels (N) int32``."""
    _need_cuda(logits)
    if logits.dtype != torch.float32:
        raise RuntimeError("softmax_argmax expects float32 logits")
    logits = logits.contiguous()
    N, K = logits.shape
    S = torch.empty_like(logits)
    labels = torch.empty(N, dtype=torch.int32, device=logits.device)
    with torch.cuda.device(logits.device):
        call("mg_softmax_argmax", logits.data_ptr(), N, K, S.data_ptr(), labels.data_ptr(), _stream())
    return S, labels


# ---------------------------------------------------------------------------------------------
# normalized cut
# ---------------------------------------------------------------------------------------------
def ncut_edge_weights(h: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    """``w_e = exp(-|h[src]-h[tgt]|^2 / 2)`` in COO order (mincut_refinement.py:43-51)."""
    _need_cuda(h, edge_index)
    if h.dtype != torch.float32:
        raise RuntimeError("ncut_edge_weights expects float32 features")
    h = h.contiguous()
    edge_index = edge_index.long().contiguous()
    N, D = h.shape
    E = edge_index.shape[1]
    w = torch.empty(E, dtype=torch.float32, device=h.device)
    with torch.cuda.device(h.device):
        call("mg_ncut_edge_weights", h.data_ptr(), N, D, edge_index.data_ptr(), E, w.data_ptr(), _stream())
    return w


def ncut_loss(h: torch.Tensor, S: torch.Tensor, rowptr_out: torch.Tensor, col_out: torch.Tensor,
              nodes_per_graph: int = 0, with_stats: bool = False):
    """Soft normalized-cut loss per graph -> ``(G,)`` float32 (mincut_refinement.py:55-160).
    ``with_stats``: also return ``(G, 2K)`` = per-graph ``[assoc | cut]`` for :func:`ncut_backward`."""
    _need_cuda(h, S, rowptr_out, col_out)
    if h.dtype != torch.float32 or S.dtype != torch.float32:
        raise RuntimeError("ncut_loss expects float32 tensors")
    h, S = h.contiguous(), S.contiguous()
    N, D = h.shape
    K = S.shape[1]
    G = N // nodes_per_graph if nodes_per_graph > 0 else 1
    loss = torch.empty(G, dtype=torch.float32, device=h.device)
    stats = torch.empty((G, 2 * K), dtype=torch.float32, device=h.device) if with_stats else None
    work = torch.empty(int(_lib.load().mg_ncut_work_bytes(N, K, G)), dtype=torch.uint8, device=h.device)
    with torch.cuda.device(h.device):
        call("mg_ncut_loss", h.data_ptr(), S.data_ptr(), rowptr_out.data_ptr(), col_out.data_ptr(), N, D, K,
             int(nodes_per_graph), loss.data_ptr(), _ptr(stats), work.data_ptr(), _stream())
    return (loss, stats) if with_stats else loss


def ncut_backward(h, S, rowptr_out, col_out, rowptr_in, col_in, stats, grad_loss, nodes_per_graph: int = 0):
    """Gradients of the per-graph N-cut loss w.r.t. ``h (N,D)`` and ``S (N,K)``."""
    _need_cuda(h, S, stats, grad_loss)
    h, S = h.contiguous(), S.contiguous()
    N, D = h.shape
    K = S.shape[1]
    gh = torch.empty_like(h)
    gS = torch.empty_like(S)
    grad_loss = grad_loss.contiguous().float()
    with torch.cuda.device(h.device):
        call("mg_ncut_backward", h.data_ptr(), S.data_ptr(), rowptr_out.data_ptr(), col_out.data_ptr(), rowptr_in.data_ptr(),
             col_in.data_ptr(), N, D, K, int(nodes_per_graph), stats.data_ptr(), grad_loss.data_ptr(), gh.data_ptr(),
             gS.data_ptr(), _stream())
    return gh, gS


# ---------------------------------------------------------------------------------------------
# backward of the glue stages (training)
# ---------------------------------------------------------------------------------------------
def unpool_nearest_backward(grad_out: torch.Tensor, labels: Optional[torch.Tensor], K: int, Hp: int, Wp: int) -> torch.Tensor:
    """``grad_table (B,K,D)`` = sum of ``grad_out (B,D,H,W)`` over the pixels whose patch carries label k
    (dual of :func:`unpool_nearest`)."""
    _need_cuda(grad_out, labels)
    if grad_out.dim() != 4:
        raise ValueError("grad_out must be (B, D, H, W)")
    if grad_out.stride()[1:] != (grad_out.shape[2] * grad_out.shape[3], grad_out.shape[3], 1):
        grad_out = grad_out.contiguous()
    B, D, H, W = grad_out.shape
    if labels is not None:
        labels = labels.contiguous()
    work = torch.empty(int(_lib.load().mg_unpool_backward_work_bytes(B, D, Hp, Wp, K)), dtype=torch.uint8, device=grad_out.device)
    gt = torch.empty((B, K, D), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        call("mg_unpool_nearest_backward", grad_out.data_ptr(), _dtype_code(grad_out.dtype),
             grad_out.stride(0) if B > 1 else D * H * W, _ptr(labels), B, K, D, Hp, Wp, H, W, work.data_ptr(), gt.data_ptr(),
             _stream())
    return gt


def segment_mean_backward(grad_out: torch.Tensor, labels: torch.Tensor, counts: torch.Tensor, N: int) -> torch.Tensor:
    """``grad_h (B,N,D) = grad_out[b, labels[b,n], :] / counts[b, labels[b,n]]`` (dual of :func:`segment_mean`)."""
    _need_cuda(grad_out, labels, counts)
    grad_out = grad_out.contiguous().float()
    B, K, D = grad_out.shape
    gh = torch.empty((B, N, D), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        call("mg_segment_mean_backward", grad_out.data_ptr(), labels.contiguous().data_ptr(), counts.contiguous().data_ptr(),
             B, N, D, K, 0, gh.data_ptr(), _stream())
    return gh


def softmax_backward(S: torch.Tensor, grad_S: torch.Tensor) -> torch.Tensor:
    _need_cuda(S, grad_S)
    S, grad_S = S.contiguous(), grad_S.contiguous().float()
    N, K = S.shape
    gl = torch.empty_like(S)
    with torch.cuda.device(S.device):
        call("mg_softmax_backward", S.data_ptr(), grad_S.data_ptr(), N, K, gl.data_ptr(), _stream())
    return gl


# ---------------------------------------------------------------------------------------------
# fused per-image block
# ---------------------------------------------------------------------------------------------
def block_supported(B: int, Hp: int, Wp: int, in_dim: int, D: int, H1: int, H2: int, H3: int, K: int) -> bool:
    return bool(_lib.load().mg_block_supported(B, Hp, Wp, in_dim, D, H1, H2, H3, K))


def block_prepare(W1, a1, W2, a2, W3, a3, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Re-arrange the three GAT nets' weights for ``block_forward`` (once per weight version).
    ``out``: refresh an existing blob in place (keeps captured CUDA graphs valid)."""
    _need_cuda(W1, a1, W2, a2, W3, a3)
    H1, D, in_dim = W1.shape
    H2, K, _ = W2.shape
    H3 = W3.shape[0]
    ws = [t.detach().contiguous().float() for t in (W1, a1, W2, a2, W3, a3)]
    n = int(_lib.load().mg_block_prep_floats(in_dim, D, H1, H2, H3, K))
    prep = out if out is not None and out.numel() == n and out.device == W1.device else \
        torch.empty(n, dtype=torch.float32, device=W1.device)
    with torch.cuda.device(W1.device):
        call("mg_block_prepare", *[w.data_ptr() for w in ws], in_dim, D, H1, H2, H3, K, prep.data_ptr(), _stream())
    return prep


def block_forward(x: torch.Tensor, Hp: int, Wp: int, prep: torch.Tensor, D: int, H1: int, H2: int, H3: int, K: int,
                  slopes=(0.2, 0.2, 0.2), want_region_in: bool = False, outs=None, peer=None, feature_loss=None):
    """One launch for the whole per-image pipeline.  ``x (B,N,in)`` f32|bf16.  Returns
    ``h (B,N,D), S (B,N,K), labels (B,N) int32, loss (B,), region_in|None, region_out (B,K,D)``.
    ``peer``: a ``_lib.PeerOut`` (``distributed.PeerExchange.slot(i)``) — the kernel then also stores
    ``loss | region_out | labels`` into every rank's gathered buffer over NVLink and publishes the step
    (``mg_block_forward_push``).  ``feature_loss``: ``(f_unet (B,N,D) f32, y (B,N), margin)`` — FeatureConsistencyLoss
    evaluated inside the kernel on the rows of ``h`` (``mg_block_forward_ex``); the per-image sums ``(B,)`` are
    appended to the returned tuple."""
    _need_cuda(x, prep)
    x = x.contiguous()
    B, N, in_dim = x.shape
    dev = x.device
    f32 = dict(dtype=torch.float32, device=dev)
    if outs is not None:
        # caller-owned (contiguous, fp32 / int32) output buffers, e.g. batch slices shared by several shards
        h, S, labels, loss, rout = outs
        if tuple(h.shape) != (B, N, D) or tuple(S.shape) != (B, N, K) or tuple(labels.shape) != (B, N) or \
                tuple(loss.shape) != (B,) or tuple(rout.shape) != (B, K, D) or \
                not all(t.is_contiguous() for t in outs) or labels.dtype != torch.int32:
            raise ValueError("block_forward: bad output buffers")
    else:
        h = torch.empty((B, N, D), **f32)
        S = torch.empty((B, N, K), **f32)
        labels = torch.empty((B, N), dtype=torch.int32, device=dev)
        loss = torch.empty(B, **f32)
        rout = torch.empty((B, K, D), **f32)
    q = torch.empty((B, N, 2 * H2 + H2 * K), **f32)
    rin = torch.empty((B, K, D), **f32) if want_region_in else None
    fl = fl_out = None
    if feature_loss is not None:
        fu, y, margin = feature_loss
        _need_cuda(fu, y)
        if fu.dtype != torch.float32 or tuple(fu.shape) != (B, N, D) or tuple(y.shape) != (B, N):
            raise ValueError("block_forward: feature_loss needs f_unet (B,N,D) float32 and y (B,N)")
        fu, y = fu.contiguous(), y.to(torch.float32).contiguous()
        fl_out = torch.empty(B, **f32)
        fl = _lib.BlockFeatureLoss(fu.data_ptr(), y.data_ptr(), float(margin), fl_out.data_ptr())
    with torch.cuda.device(dev):
        call("mg_block_forward_ex", x.data_ptr(), _dtype_code(x.dtype), B, Hp, Wp, in_dim, D, H1, H2, H3, K,
             float(slopes[0]), float(slopes[1]), float(slopes[2]), prep.data_ptr(), h.data_ptr(), q.data_ptr(),
             S.data_ptr(), labels.data_ptr(), loss.data_ptr(), _ptr(rin), rout.data_ptr(),
             None if peer is None else C.addressof(peer), None if fl is None else C.addressof(fl), _stream())
    if feature_loss is not None:
        return h, S, labels, loss, rin, rout, fl_out
    return h, S, labels, loss, rin, rout


# ---------------------------------------------------------------------------------------------
# losses on tensors the block already holds (scope row f4)
# ---------------------------------------------------------------------------------------------
_YDT = {torch.float32: MG_F32, torch.bfloat16: MG_BF16, torch.int32: _lib.MG_I32, torch.int64: _lib.MG_I64}


def _label_tensor(y: torch.Tensor) -> torch.Tensor:
    """Labels in a dtype the kernels read directly; other dtypes (bool, uint8, fp16, ...) take the
    reference's own ``.float()`` (feature_loss.py:106)."""
    return y.contiguous() if y.dtype in _YDT else y.float().contiguous()


def feature_consistency_loss(f_unet: torch.Tensor, f_graph: torch.Tensor, y: torch.Tensor, margin: float = 1.0,
                             with_per_image: bool = False):
    """``mean_b sum_n [y*d^2 + (1-y)*max(0, margin-d)^2]``, ``d = |f_unet - f_graph|`` per patch
    (model/unet/feature_loss.py:103-123).  ``f_* (B,N,D)`` f32|bf16, ``y (B,N)``.  Returns a 0-dim
    float32 tensor (and the ``(B,)`` per-image sums)."""
    dev = _need_cuda(f_unet, f_graph, y)
    f_unet, f_graph, y = f_unet.contiguous(), f_graph.contiguous(), _label_tensor(y)
    B, N, D = f_unet.shape
    loss = torch.empty((), dtype=torch.float32, device=dev)
    per = torch.empty(B, dtype=torch.float32, device=dev) if with_per_image else None
    work = torch.empty(int(_lib.load().mg_feature_loss_work_bytes(B, N)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        call("mg_feature_consistency_loss", f_unet.data_ptr(), _dtype_code(f_unet.dtype), f_graph.data_ptr(),
             _dtype_code(f_graph.dtype), y.data_ptr(), _YDT[y.dtype], B, N, D, float(margin), work.data_ptr(), _ptr(per),
             loss.data_ptr(), _stream())
    return (loss, per) if with_per_image else loss


def feature_consistency_loss_backward(f_unet, f_graph, y, margin: float, grad_loss: torch.Tensor, need_unet: bool = True,
                                      need_graph: bool = True):
    """float32 gradients w.r.t. ``f_unet`` and ``f_graph`` (``None`` where not needed)."""
    dev = _need_cuda(f_unet, f_graph, y, grad_loss)
    f_unet, f_graph, y = f_unet.contiguous(), f_graph.contiguous(), _label_tensor(y)
    B, N, D = f_unet.shape
    gu = torch.empty((B, N, D), dtype=torch.float32, device=dev) if need_unet else None
    gg = torch.empty((B, N, D), dtype=torch.float32, device=dev) if need_graph else None
    grad_loss = grad_loss.contiguous().float()
    with torch.cuda.device(dev):
        call("mg_feature_consistency_loss_backward", f_unet.data_ptr(), _dtype_code(f_unet.dtype), f_graph.data_ptr(),
             _dtype_code(f_graph.dtype), y.data_ptr(), _YDT[y.dtype], B, N, D, float(margin), grad_loss.data_ptr(),
             _ptr(gu), _ptr(gg), _stream())
    return gu, gg


def tv_loss(x: torch.Tensor, weight: float = 1.0, with_terms: bool = False):
    """``weight * (h_tv/((H-1)W) + w_tv/(H(W-1))) / B`` over ``x (B,C,H,W)`` f32|bf16
    (scripts/train_end_to_end.py:73-89).  Returns a 0-dim float32 tensor (view of a 3-float result
    ``[loss, h_tv, w_tv]`` when ``with_terms``)."""
    dev = _need_cuda(x)
    if x.dim() != 4:
        raise RuntimeError("tv_loss expects (B, C, H, W)")
    x = x.contiguous()
    B, C, H, W = x.shape
    code = _dtype_code(x.dtype)
    out3 = torch.empty(3, dtype=torch.float32, device=dev)
    work = torch.empty(int(_lib.load().mg_tv_loss_work_bytes(code, B, C, H, W)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        call("mg_tv_loss", x.data_ptr(), code, B, C, H, W, float(weight), work.data_ptr(), out3.data_ptr(), _stream())
    return out3 if with_terms else out3[0]


def tv_loss_backward(x: torch.Tensor, weight: float, grad_loss: torch.Tensor) -> torch.Tensor:
    dev = _need_cuda(x, grad_loss)
    x = x.contiguous()
    B, C, H, W = x.shape
    gx = torch.empty((B, C, H, W), dtype=torch.float32, device=dev)
    grad_loss = grad_loss.contiguous().float()
    with torch.cuda.device(dev):
        call("mg_tv_loss_backward", x.data_ptr(), _dtype_code(x.dtype), B, C, H, W, float(weight), grad_loss.data_ptr(),
             gx.data_ptr(), _stream())
    return gx


# ---------------------------------------------------------------------------------------------
# fusion: per-region embeddings -> dense per-pixel map (csrc/fusion.cu) — compiled, not yet run on hardware
# ---------------------------------------------------------------------------------------------
def region_map_gather(table: torch.Tensor, region_map: torch.Tensor, out: Optional[torch.Tensor] = None,
                      out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """``out[b,d,y,x] = table[region_map[b,y,x], d]`` where the index is in ``[0, R)``, else 0
    (FeatureFusion.forward per-region branch, model/fusion_detection/feature_fusion.py:81-140).
    ``table (R,D)`` float32, ``region_map (B,H,W)`` int32 | int64; ``out`` may be a channel slice
    ``buf[:, c0:c0+D]`` of a contiguous ``(B,Ctot,H,W)`` buffer."""
    _need_cuda(table, region_map, out)
    if table.dim() != 2 or region_map.dim() != 3:
        raise ValueError("region_map_gather expects table (R,D) and region_map (B,H,W)")
    if table.dtype != torch.float32:
        raise RuntimeError("region_map_gather expects a float32 table")
    if region_map.dtype not in (torch.int32, torch.int64):
        raise RuntimeError("region_map must be int32 or int64")
    table, region_map = table.contiguous(), region_map.contiguous()
    R, D = table.shape
    B, H, W = region_map.shape
    if R == 0 or D == 0 or B * H * W == 0:
        raise ValueError("region_map_gather: empty table or map")
    if out is None:
        out = torch.empty((B, D, H, W), dtype=out_dtype, device=table.device)
    if tuple(out.shape) != (B, D, H, W) or out.stride()[1:] != (H * W, W, 1):
        raise ValueError("out must be (B,D,H,W) with contiguous (D,H,W) planes")
    with torch.cuda.device(table.device):
        call("mg_region_map_gather", table.data_ptr(), R, D, region_map.data_ptr(),
             _lib.MG_I32 if region_map.dtype == torch.int32 else _lib.MG_I64, B, H, W, out.data_ptr(),
             _dtype_code(out.dtype), out.stride(0) if B > 1 else D * H * W, _stream())
    return out


# ---------------------------------------------------------------------------------------------
# peer-memory exchange (multi-GPU; csrc/peer_exchange.cu, the push itself is inside block_forward)
# ---------------------------------------------------------------------------------------------
def peer_wait(flags: torch.Tensor, first_flag: int, world: int, seq: torch.Tensor, status: Optional[torch.Tensor] = None) -> None:
    """Make the current stream wait until ``flags[first_flag + r] >= seq[0]`` for every source rank ``r`` (``seq``: the
    slot's device-side step counter, advanced by the block kernel).  Bounded spin; ``status[0] = 1`` on expiry."""
    dev = _need_cuda(flags, seq, status)
    with torch.cuda.device(dev):
        call("mg_peer_wait", flags.data_ptr(), int(first_flag), int(world), seq.data_ptr(), _ptr(status), _stream())


def peer_mem_alloc(nbytes: int, device) -> Tuple[int, bytes]:
    """cudaMalloc + zero + CUDA IPC handle on ``device``: ``(device pointer, 64-byte handle)``.  Set-up call (synchronises)."""
    ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
    with torch.cuda.device(device):
        call("mg_peer_mem_alloc", int(nbytes), C.addressof(ptr), C.addressof(handle))
    return int(ptr.value), bytes(handle)


def peer_mem_open(handle: bytes, device) -> int:
    """Map another process's ``peer_mem_alloc`` allocation into this process; returns the device pointer."""
    ptr, h = C.c_void_p(), (C.c_ubyte * 64).from_buffer_copy(handle)
    with torch.cuda.device(device):
        call("mg_peer_mem_open", C.addressof(h), C.addressof(ptr))
    return int(ptr.value)


def peer_mem_close(ptr: int) -> None:
    call("mg_peer_mem_close", int(ptr))


def peer_mem_free(ptr: int) -> None:
    call("mg_peer_mem_free", int(ptr))


class _DeviceMemory:
    """``__cuda_array_interface__`` view of a raw allocation, so torch can wrap it without owning it."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def tensor_from_ptr(ptr: int, nbytes: int, device) -> torch.Tensor:
    """uint8 tensor over ``[ptr, ptr + nbytes)`` on ``device`` (memory stays owned by the caller)."""
    with torch.cuda.device(device):
        return torch.as_tensor(_DeviceMemory(ptr, nbytes), device=torch.device(device))
