"""FP32 CPU restatement of the MinGraph-UNet graph block (TEST INFRASTRUCTURE).

Every function states the reference lines it follows (paths are relative to
``/root/reference/MinGraph-UNet``).  The restatement keeps the reference's
*operation order* (dense transform -> per-edge gathers -> global-max shifted exp
-> scatter-add by target) so that (a) its fp32 rounding behaviour is the
reference's, and (b) timing it is a fair stand-in for the reference's CPU path
(``bench.py`` ``cpu_baseline.kind == "port"``).

It is written as stateless functions over plain tensors; weights are passed as
``Ws`` ``(H, F, in)`` and ``As`` ``(H, 2F)`` stacks (the reference keeps one
``nn.Linear`` pair per head: ``gat_layers.0.heads.{h}.W.weight`` / ``.a.weight``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# graph construction
# ----------------------------------------------------------------------------
def grid_dims(H: int, W: int, patch: int = 16) -> Tuple[int, int]:
    """ceil-div patch grid.  preprocessing/graph_construction/patch_graph_construction.py:67-68"""
    return (H + patch - 1) // patch, (W + patch - 1) // patch


def grid_edge_index_loop(Hp: int, Wp: int) -> np.ndarray:
    """4-connectivity COO edge list in the reference's emission order.

    patch_graph_construction.py:78-97 (row-major node walk; right pair, then
    down pair; ``(2,0)`` for a single patch).  Pure-Python loop: small cases only.
    """
    src, tgt = [], []
    for r in range(Hp):
        for c in range(Wp):
            n = r * Wp + c
            if c + 1 < Wp:
                src += [n, n + 1]
                tgt += [n + 1, n]
            if r + 1 < Hp:
                src += [n, n + Wp]
                tgt += [n + Wp, n]
    return np.asarray([src, tgt], dtype=np.int64).reshape(2, -1)


def grid_edge_index(Hp: int, Wp: int) -> np.ndarray:
    """Vectorised numpy form of :func:`grid_edge_index_loop` (same order, int64)."""
    r, c = np.divmod(np.arange(Hp * Wp, dtype=np.int64), Wp)
    n = r * Wp + c
    has_r = c + 1 < Wp
    has_d = r + 1 < Hp
    # each node emits up to 4 slots: (n,n+1) (n+1,n) (n,n+Wp) (n+Wp,n)
    s = np.stack([n, n + 1, n, n + Wp], 1)
    t = np.stack([n + 1, n, n + Wp, n], 1)
    keep = np.stack([has_r, has_r, has_d, has_d], 1)
    return np.stack([s[keep], t[keep]], 0).astype(np.int64)


def complete_edge_index(K: int) -> np.ndarray:
    """Complete digraph on K regions.  scripts/train_end_to_end.py:376-380."""
    if K <= 1:
        return np.zeros((2, 0), dtype=np.int64)
    s, t = np.triu_indices(K, 1)
    return np.stack([np.concatenate([s, t]), np.concatenate([t, s])], 0).astype(np.int64)


# ----------------------------------------------------------------------------
# patches / pooling
# ----------------------------------------------------------------------------
def image_to_patches(img: torch.Tensor, patch: int = 16):
    """patch_graph_construction.py:27-47 — zero pad right/bottom, unfold twice,
    ``permute(1,2,0,3,4)`` -> ``(N, C, P, P)`` plus ``(nph, npw)``."""
    C, H, W = img.shape
    ph = (patch - H % patch) % patch
    pw = (patch - W % patch) % patch
    if ph or pw:
        img = F.pad(img, (0, pw, 0, ph))
    t = img.unfold(1, patch, patch).unfold(2, patch, patch)
    nph, npw = t.shape[1], t.shape[2]
    t = t.permute(1, 2, 0, 3, 4).contiguous().view(-1, C, patch, patch)
    return t, (nph, npw)


def patch_mean_pool(fmap: torch.Tensor, patch: int = 16) -> torch.Tensor:
    """Per-channel patch mean ``(C,Hf,Wf) -> (N,C)``: the pooling the reference
    documents (patch_graph_construction.py:104-109, "averaging U-Net encoder
    features within the patch boundaries") expressed with its own patch
    extractor, i.e. ``image_to_patches(fmap)[0].mean(dim=(2,3))``.  Zero padding
    counts in the mean (divisor is always P*P) exactly as ``patches.mean`` does in
    scripts/graph_refinement.py:78,98,103."""
    p, _ = image_to_patches(fmap, patch)
    return p.mean(dim=(2, 3))


# ----------------------------------------------------------------------------
# GAT
# ----------------------------------------------------------------------------
def gat_head(x, ei, W, a, alpha: float = 0.2):
    """One attention head, eval mode.  model/gat/graph_attention.py:53-118."""
    N = x.shape[0]
    Fo = W.shape[0]
    Wh = x @ W.t()                                             # :53
    s = Wh[ei[0]]                                              # :57
    t = Wh[ei[1]]                                              # :58
    e = F.leaky_relu(torch.cat([s, t], 1) @ a.view(1, -1).t(), alpha)   # :61-65
    ex = torch.exp(e - torch.max(e))                           # :86  (global max)
    den = torch.zeros(N, 1).scatter_add_(0, ei[1].unsqueeze(1), ex)     # :90-91
    att = ex / (den[ei[1]] + 1e-10)                            # :94-96
    hp = torch.zeros(N, Fo)
    hp.scatter_add_(0, ei[1].unsqueeze(1).repeat(1, Fo), att * s)       # :104-112
    return F.elu(hp)                                           # :118


def gat_layer(x, ei, Ws, As, alpha: float = 0.2, concat: bool = False):
    """Multi-head layer, eval mode (dropout off).  graph_attention.py:150-160."""
    outs = [gat_head(x, ei, Ws[h], As[h], alpha) for h in range(Ws.shape[0])]
    if concat:
        return torch.cat(outs, 1)
    return torch.mean(torch.stack(outs, 0), 0)


def gat_network(x, ei, Ws, As, alpha: float = 0.2):
    """1-layer ``GATNetwork`` = one averaging multi-head layer.  graph_attention.py:168-172,188-192."""
    return gat_layer(x, ei, Ws, As, alpha, concat=False)


def gat_network_multilayer(x, ei, layers, alpha: float = 0.2):
    """Working multi-layer stack (scope row f4): ``layers`` = list of ``(Ws, As)``; every layer but
    the last concatenates its heads (graph_attention.py:137-139,155,176,181), the last averages
    (:141,158,185); eval mode, so the inter-layer ``nn.Dropout`` (:160) is the identity.  This is
    the composition of reference ``MultiHeadGATLayer``s with the widths the concatenating layer
    really emits (the reference's own ``GATNetwork`` wires ``hidden*heads`` there and crashes)."""
    h = x
    for i, (Ws, As) in enumerate(layers):
        h = gat_layer(h, ei, Ws, As, alpha, concat=i + 1 < len(layers))
    return h


# ----------------------------------------------------------------------------
# normalized cut
# ----------------------------------------------------------------------------
def ncut_edge_weights(h, ei):
    """model/graph_partition/mincut_refinement.py:43-51 (sigma fixed at 1.0)."""
    d = torch.sum((h[ei[0]] - h[ei[1]]) ** 2, dim=1)
    return torch.exp(-d / 2.0)


def ncut_loss(h, ei, S, K: int):
    """mincut_refinement.py:77-160.  Returns a 0-dim tensor (0.0 if no segment
    has association > 1e-8, where the reference returns a Python float)."""
    N = h.shape[0]
    if tuple(S.shape) != (N, K):
        raise ValueError("segment_assignments_soft shape mismatch.")      # :73-74
    w = ncut_edge_weights(h, ei)
    total = torch.zeros(())
    for k in range(K):
        p = S[:, k]
        deg = torch.zeros(N).scatter_add_(0, ei[0], w)                    # :92-96 (by SOURCE)
        assoc = torch.sum(p * deg)                                       # :102
        cut = torch.sum(w * p[ei[0]] * (1 - p[ei[1]]))                   # :112-113,149
        if assoc > 1e-8:                                                 # :151-152
            total = total + cut / assoc
    return total


def mincut_forward(h, ei, K: int, predictor):
    """mincut_refinement.py:192-205: logits -> softmax -> (loss, S)."""
    S = F.softmax(predictor(h, ei), dim=1)
    return ncut_loss(h, ei, S, K), S


# ----------------------------------------------------------------------------
# region stage + un-pool
# ----------------------------------------------------------------------------
def region_mean_pool(h, hard, K: int):
    """scripts/train_end_to_end.py:368-373 — masked mean per label, 0 if empty."""
    out = torch.zeros(K, h.shape[1])
    for k in range(K):
        m = hard == k
        if m.sum() > 0:
            out[k] = h[m].mean(dim=0)
    return out


def unpool_nearest(fpatch, nph: int, npw: int, H: int, W: int):
    """train_end_to_end.py:411-421 — ``(N,D)`` -> ``(D,nph,npw)`` -> nearest
    up-sampling to ``(D,H,W)``."""
    D = fpatch.shape[1]
    g = fpatch.t().reshape(D, nph, npw)
    return F.interpolate(g.unsqueeze(0), size=(H, W), mode="nearest").squeeze(0)


def nearest_index(out_size: int, in_size: int) -> np.ndarray:
    """Source index table of torch's CPU ``upsample_nearest`` (legacy 'nearest'):
    identity / ``>>1`` special cases, otherwise ``min(floor(dst * float32(in/out)), in-1)``
    evaluated in float32.  Used by tests to pin the un-pool index math."""
    d = np.arange(out_size, dtype=np.int64)
    if out_size == in_size:
        return d
    if out_size == 2 * in_size:
        return d >> 1
    scale = np.float32(in_size) / np.float32(out_size)
    idx = np.floor(d.astype(np.float32) * scale).astype(np.int64)
    return np.minimum(idx, in_size - 1)


def graph_block_image(x, H: int, W: int, params: Dict[str, torch.Tensor], K: int = 2,
                      patch: int = 16, alpha: float = 0.2, want_dense: bool = True, hard=None):
    """One image through the block, stage order of train_end_to_end.py:318-421
    (node features passed in instead of ``randn`` :326; the crashing feature-loss
    call :344 omitted).  ``params`` holds ``patch_W/patch_a``, ``pred_W/pred_a``,
    ``region_W/region_a`` head stacks.  Returns a dict of every intermediate.
    ``hard``: take these patch labels instead of ``argmax(S)`` for everything downstream of :356 (the
    parity tests re-derive the tail from the labels the device produced, so an argmax flip at a
    numerical tie cannot hide the region / un-pool stages from the comparison)."""
    nph, npw = grid_dims(H, W, patch)
    if x.shape[0] != nph * npw:                                           # patch_graph_construction.py:71-74
        raise ValueError("patch feature count does not match the patch grid")
    ei = torch.from_numpy(grid_edge_index(nph, npw))                     # :329
    h = gat_network(x, ei, params["patch_W"], params["patch_a"], alpha)  # :332
    if ei.shape[1] == 0:
        raise RuntimeError("empty edge_index")
    logits = gat_network(h, ei, params["pred_W"], params["pred_a"], alpha)
    S = F.softmax(logits, dim=1)                                          # mincut_refinement.py:193
    loss = ncut_loss(h, ei, S, K)                                         # :196
    hard_own = torch.argmax(S, dim=1)                                     # train_end_to_end.py:356
    hard = hard_own if hard is None else torch.as_tensor(hard).long()
    R = region_mean_pool(h, hard, K)                                      # :368-373
    rei = torch.from_numpy(complete_edge_index(K))                       # :376-380
    if K > 0 and rei.numel() > 0:                                         # :383-389
        G = gat_network(R, rei, params["region_W"], params["region_a"], alpha)
    else:
        G = R
    P = G[hard]                                                           # :403-406
    out = dict(edge_index=ei, h=h, logits=logits, S=S, loss=loss, hard=hard, hard_argmax=hard_own,
               region_in=R, region_out=G, f_patch=P, grid=(nph, npw))
    if want_dense:
        out["f_g"] = unpool_nearest(P, nph, npw, H, W)                    # :411-421
    return out


# ----------------------------------------------------------------------------
# kNN graph (NOT in the reference — parity unpinned by it; this restatement IS the definition)
# ----------------------------------------------------------------------------
def knn_sqdist(x: np.ndarray) -> np.ndarray:
    """All-pairs squared Euclidean distances, fp32, accumulated in FEATURE ORDER with separately rounded
    subtract / multiply / add (numpy float32 ops never fuse), so a CUDA kernel that uses
    ``__fsub_rn/__fmul_rn/__fadd_rn`` in the same order reproduces every bit.  The nearest reference
    arithmetic is the per-edge ``sum((f_src - f_tgt)**2)`` of mincut_refinement.py:43-46."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = x.shape[0]
    acc = np.zeros((n, n), dtype=np.float32)
    for d in range(x.shape[1]):
        t = x[:, None, d] - x[None, :, d]
        acc = acc + t * t
    return acc


def knn_graph(x: np.ndarray, k: int, nodes_per_graph: int = 0):
    """k nearest OTHER nodes of every node within its graph; ties to the lower node id; neighbours sorted by
    (distance, id).  Returns ``edge_index (2, N*k) int64`` (row 0 = neighbour = source, row 1 = node = target)
    and the distances ``(N, k) float32``."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    N = x.shape[0]
    npg = nodes_per_graph if nodes_per_graph > 0 else N
    if N % npg or npg <= k:
        raise ValueError("bad graph size for kNN")
    src = np.empty((N, k), dtype=np.int64)
    dist = np.empty((N, k), dtype=np.float32)
    for g0 in range(0, N, npg):
        d = knn_sqdist(x[g0:g0 + npg])
        np.fill_diagonal(d, np.inf)                                   # self excluded
        ids = np.arange(npg)
        order = np.lexsort((np.broadcast_to(ids, d.shape), d), axis=1)[:, :k]     # primary key d, then id
        src[g0:g0 + npg] = order + g0
        dist[g0:g0 + npg] = np.take_along_axis(d, order, 1)
    tgt = np.repeat(np.arange(N, dtype=np.int64), k)
    return np.stack([src.reshape(-1), tgt]), dist


# ----------------------------------------------------------------------------
# weight helpers (reference init, graph_attention.py:36-37)
# ----------------------------------------------------------------------------
def init_gat_params(in_dim: int, out_dim: int, heads: int, gen: Optional[torch.Generator] = None):
    """xavier_uniform(gain=1.414) head stacks ``(H,F,in)``, ``(H,2F)``."""
    Ws = torch.empty(heads, out_dim, in_dim)
    As = torch.empty(heads, 2 * out_dim)
    for h in range(heads):
        bw = 1.414 * math.sqrt(6.0 / (in_dim + out_dim))
        ba = 1.414 * math.sqrt(6.0 / (2 * out_dim + 1))
        Ws[h].uniform_(-bw, bw, generator=gen)
        As[h].uniform_(-ba, ba, generator=gen)
    return Ws, As


def init_block_params(in_dim: int = 20, out_dim: int = 64, heads: int = 4, K: int = 2, seed: int = 1234):
    """The three GAT nets the reference builds (train_end_to_end.py:144-178):
    patch ``in->out`` (heads), predictor ``out->K`` (max(1,heads//2)), region ``out->out`` (heads)."""
    g = torch.Generator().manual_seed(seed)
    pw, pa = init_gat_params(in_dim, out_dim, heads, g)
    qw, qa = init_gat_params(out_dim, K, max(1, heads // 2), g)
    rw, ra = init_gat_params(out_dim, out_dim, heads, g)
    return dict(patch_W=pw, patch_a=pa, pred_W=qw, pred_a=qa, region_W=rw, region_a=ra)


def stack_from_state_dict(sd, prefix: str = "gat_layers.0.heads."):
    """Collect ``(H,F,in)`` / ``(H,2F)`` stacks from a reference ``GATNetwork`` state_dict."""
    hs = sorted({int(k[len(prefix):].split(".")[0]) for k in sd if k.startswith(prefix)})
    Ws = torch.stack([sd[f"{prefix}{h}.W.weight"] for h in hs], 0).float()
    As = torch.stack([sd[f"{prefix}{h}.a.weight"].reshape(-1) for h in hs], 0).float()
    return Ws, As


# ----------------------------------------------------------------------------
# losses on block tensors (scope row f4)
# ----------------------------------------------------------------------------
def feature_consistency_loss(f_unet, f_graph, y, margin: float = 1.0):
    """model/unet/feature_loss.py:103-123 with the revised-argument reading the code implements:
    ``f_* (B,N,D)``, ``y (B,N)``; positive term ``y*dist_sq`` (:110), ``dist = sqrt(dist_sq + 1e-8)``
    (:116), hinge ``relu(margin - dist)`` (:118), negative term ``(1-y)*hinge**2`` (:119), sum over
    patches then mean over the batch (:124)."""
    y_p = y.float().unsqueeze(-1)
    dist_sq = torch.sum((f_unet - f_graph) ** 2, dim=2)
    loss_positive = y_p.squeeze(-1) * dist_sq
    dist = torch.sqrt(dist_sq + 1e-8)
    hinge_term = F.relu(margin - dist)
    loss_negative = (1 - y_p.squeeze(-1)) * (hinge_term ** 2)
    return torch.sum(loss_positive + loss_negative, dim=1).mean()


def tv_loss(x, weight: float = 1.0):
    """scripts/train_end_to_end.py:79-89 — squared forward differences along H and W, each normalised
    by its element count, scaled by ``weight`` and divided by the batch size."""
    batch_size, h_x, w_x = x.size(0), x.size(2), x.size(3)
    count_h = (h_x - 1) * w_x
    count_w = h_x * (w_x - 1)
    h_tv = torch.pow(x[:, :, 1:, :] - x[:, :, :-1, :], 2).sum()
    w_tv = torch.pow(x[:, :, :, 1:] - x[:, :, :, :-1], 2).sum()
    return weight * (h_tv / count_h + w_tv / count_w) / batch_size


# ----------------------------------------------------------------------------
# feature fusion (the consumer of the block's output; scope row f1)
# ----------------------------------------------------------------------------
def region_map_gather(f_g: torch.Tensor, region_to_pixel_map: torch.Tensor) -> torch.Tensor:
    """model/fusion_detection/feature_fusion.py:81-132: ``f_g (R, D)`` + ``region_to_pixel_map (B, H, W)`` ->
    ``(B, D, H, W)`` float32; pixels whose index is outside ``[0, R)`` (``valid_mask``, :119) stay zero (:85)."""
    B, H, W = region_to_pixel_map.shape
    R, D = f_g.shape
    idx = region_to_pixel_map.reshape(B, H * W).long()                      # (:115)
    valid = (idx >= 0) & (idx < R)                                          # (:119)
    rows = f_g.float()[idx.clamp(0, R - 1)]                                 # (B, H*W, D), (:124)
    rows = torch.where(valid.unsqueeze(-1), rows, torch.zeros((), dtype=torch.float32))
    return rows.permute(0, 2, 1).reshape(B, D, H, W).contiguous()          # (:131-133)


def feature_fusion(f_u_list: Sequence[torch.Tensor], f_g: torch.Tensor, target_spatial_size=None,
                   region_to_pixel_map: Optional[torch.Tensor] = None, fusion_method: str = "concat") -> torch.Tensor:
    """``FeatureFusion.forward`` (feature_fusion.py:43-153): bilinear resize of the U-Net scales (:67-73), ``cat``
    (:75), per-region gather (:81-132) or 4-D pass-through / resize (:134-138), then ``concat`` (:143) or ``add``
    (:144-148)."""
    if target_spatial_size is None:
        target_spatial_size = (f_u_list[0].size(2), f_u_list[0].size(3))
    size = (int(target_spatial_size[0]), int(target_spatial_size[1]))
    fu = [t if (t.size(2), t.size(3)) == size else F.interpolate(t, size=size, mode="bilinear", align_corners=False)
          for t in f_u_list]
    f_u = torch.cat(fu, dim=1)
    if f_g.ndim == 2 and region_to_pixel_map is not None:
        g = region_map_gather(f_g, region_to_pixel_map)
    elif f_g.ndim == 4:
        g = f_g if (f_g.size(2), f_g.size(3)) == size else F.interpolate(f_g, size=size, mode="bilinear", align_corners=False)
    else:
        raise ValueError(f"f_g has unsupported shape {f_g.shape}. "
                         "Expected (Num_regions, D_gat) with region_map or (B, D_gat, H, W).")
    if fusion_method == "concat":
        return torch.cat([f_u, g], dim=1)
    if fusion_method == "add":
        if f_u.shape[1] != g.shape[1]:
            raise ValueError("Channel dimensions must match for 'add' fusion or implement adaptation.")
        return f_u + g
    raise NotImplementedError(f"Fusion method '{fusion_method}' not implemented.")
