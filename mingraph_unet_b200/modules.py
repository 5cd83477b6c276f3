"""Drop-in ``nn.Module`` mirror of the reference's graph-block classes, executing on the sm_100a
kernels of ``libmingraph_b200.so``.

Constructor arguments, ``forward`` signatures, tensor layouts, parameter names / shapes
(``state_dict`` keys) and error behaviour follow the reference (paths relative to the
reference root):

* ``GraphAttentionLayer``   — model/gat/graph_attention.py:5-118
* ``MultiHeadGATLayer``     — model/gat/graph_attention.py:120-160
* ``GATNetwork``            — model/gat/graph_attention.py:162-192
* ``PatchGraphConstructor`` — preprocessing/graph_construction/patch_graph_construction.py:5-136
* ``MinCutRefinement``      — model/graph_partition/mincut_refinement.py:5-205
* ``PatchSegmentPredictor`` — scripts/train_end_to_end.py:40-70
* ``FeatureConsistencyLoss`` — model/unet/feature_loss.py:5-123   (scope row f4)
* ``TVLoss``                — scripts/train_end_to_end.py:73-89    (scope row f4)

so a reference checkpoint loads with ``load_state_dict`` unchanged and the classes can replace
the reference's at its two call sites (scripts/train_end_to_end.py:318-421,
scripts/graph_refinement.py:70-152).  ``GraphBlock`` (block.py) is the batched form of that
per-image loop.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .autograd import (edge_weights_apply, feature_loss_apply, gat_layer_apply, ncut_loss_apply, softmax_rows, stack_heads,
                       tv_loss_apply)
from .graph import Graph, register


def _pad_in_dim(x: torch.Tensor, W: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Kernels take in<=32, even in<=64 or in%4==0 (<=512).  Other widths are zero padded
    (exact: the padded columns contribute 0 to every dot product)."""
    in_dim = x.shape[1]
    if in_dim <= 32 or (in_dim <= 64 and in_dim % 2 == 0) or (in_dim % 4 == 0 and in_dim <= 512):
        return x, W
    if in_dim > 512:
        raise RuntimeError(f"node feature width {in_dim} > 512 is not supported by the GAT kernels")
    pad = (-in_dim) % 4
    return F.pad(x, (0, pad)), F.pad(W, (0, pad))


class GraphAttentionLayer(nn.Module):
    """One attention head (graph_attention.py:5-118).  Parameters: ``W.weight (F,in)``,
    ``a.weight (1,2F)``, xavier-uniform with gain 1.414 (:36-37)."""

    def __init__(self, in_features, out_features, dropout_rate, alpha, concat=True):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.dropout_rate = dropout_rate
        self.alpha = alpha
        self.concat = concat
        self.W = nn.Linear(in_features, out_features, bias=False)
        self.a = nn.Linear(2 * out_features, 1, bias=False)
        self.leakyrelu = nn.LeakyReLU(self.alpha)
        self.dropout = nn.Dropout(self.dropout_rate)
        nn.init.xavier_uniform_(self.W.weight, gain=1.414)
        nn.init.xavier_uniform_(self.a.weight, gain=1.414)

    def forward(self, node_features, edge_index):
        return _layer_forward([self], node_features, edge_index, concat=True, alpha=self.alpha,
                              att_dropout=self.dropout_rate if self.training else 0.0)


def _layer_forward(heads: List[GraphAttentionLayer], x: torch.Tensor, edge_index, concat: bool, alpha: float,
                   att_dropout: float, out_dtype=None) -> torch.Tensor:
    if isinstance(edge_index, Graph):
        g = edge_index
    else:
        g = Graph.from_edge_index(edge_index, x.shape[0])
    if g.E == 0:
        # graph_attention.py:86 — torch.max over an empty edge tensor raises RuntimeError
        raise RuntimeError("max(): Expected reduction dim to be specified for input.numel() == 0 "
                           "(edge_index is empty)")
    if x.dim() != 2 or x.shape[1] != heads[0].in_features:
        raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({x.shape[0]}x{x.shape[-1]} and "
                           f"{heads[0].in_features}x{heads[0].out_features})")
    if len(heads) == 1:
        W = heads[0].W.weight.unsqueeze(0)
        a = heads[0].a.weight.view(1, -1)
    else:
        W = stack_heads([h.W.weight for h in heads], heads[0].W.weight.shape)
        a = stack_heads([h.a.weight for h in heads], (heads[0].a.weight.numel(),))
    xp, Wp = _pad_in_dim(x, W)
    return gat_layer_apply(xp, g, Wp, a, concat, alpha, att_dropout, out_dtype)


class MultiHeadGATLayer(nn.Module):
    """``num_heads`` independent heads, concatenated or averaged, then dropout
    (graph_attention.py:120-160).  All heads run in ONE fused kernel sequence."""

    def __init__(self, in_features, out_features, num_heads, dropout_rate, alpha, concat=True):
        super().__init__()
        self.num_heads = num_heads
        self.concat = concat
        if concat:
            assert out_features % num_heads == 0, "out_features must be divisible by num_heads if concatenating"
            self.head_out_features = out_features // num_heads
        else:
            self.head_out_features = out_features
        self.heads = nn.ModuleList()
        for _ in range(num_heads):
            self.heads.append(GraphAttentionLayer(in_features, self.head_out_features, dropout_rate, alpha))
        self.dropout = nn.Dropout(dropout_rate)
        self.alpha = alpha
        self.dropout_rate = dropout_rate

    def forward(self, node_features, edge_index):
        out = _layer_forward(list(self.heads), node_features, edge_index, self.concat, self.alpha,
                             att_dropout=self.dropout_rate if self.training else 0.0)
        return self.dropout(out)


class GATNetwork(nn.Module):
    """Layer stack (graph_attention.py:162-192).  One layer = a single averaging multi-head layer
    (``hidden_dim`` unused).  Two or more layers are CONSTRUCTED exactly like the reference (same
    state_dict shapes), which means they inherit its width mismatch (:176-186: the next layer
    expects ``hidden_dim*num_heads`` inputs but the concatenating layer emits ``hidden_dim``) and
    raise the same RuntimeError at forward.  :class:`StackedGATNetwork` is the working stack."""

    _consistent_widths = False

    def __init__(self, node_feature_dim, hidden_dim, output_dim, num_heads, num_gat_layers=1, dropout_rate=0.1,
                 alpha=0.2):
        super().__init__()
        self.num_gat_layers = num_gat_layers
        self.gat_layers = nn.ModuleList()
        if num_gat_layers == 1:
            self.gat_layers.append(
                MultiHeadGATLayer(node_feature_dim, output_dim, num_heads, dropout_rate, alpha, concat=False))
        else:
            mid_in = hidden_dim if self._consistent_widths else hidden_dim * num_heads
            self.gat_layers.append(
                MultiHeadGATLayer(node_feature_dim, hidden_dim, num_heads, dropout_rate, alpha, concat=True))
            for _ in range(num_gat_layers - 2):
                self.gat_layers.append(
                    MultiHeadGATLayer(mid_in, hidden_dim, num_heads, dropout_rate, alpha, concat=True))
            self.gat_layers.append(
                MultiHeadGATLayer(mid_in, output_dim, num_heads, dropout_rate, alpha, concat=False))

    def forward(self, node_features, edge_index):
        h = node_features
        if not isinstance(edge_index, Graph):
            edge_index = Graph.from_edge_index(edge_index, h.shape[0])
        for layer in self.gat_layers:
            h = layer(h, edge_index)
        return h


class StackedGATNetwork(GATNetwork):
    """The WORKING multi-layer stack (scope row f4; a deliberate, flagged deviation from
    graph_attention.py:176-186): same constructor signature as :class:`GATNetwork`, but every layer
    after the first takes the ``hidden_dim`` columns the concatenating layer actually emits
    (:137-139,155), i.e. exactly the composition ``MultiHeadGATLayer(in, hidden, H, concat=True) ->
    ... -> MultiHeadGATLayer(hidden, out, H, concat=False)`` of untouched reference layers.  Its
    state_dict differs from the reference's (unusable) one only in the ``W.weight`` widths of layers
    >= 1; with ``num_gat_layers=1`` it is identical to :class:`GATNetwork`."""

    _consistent_widths = True


class PatchGraphConstructor:
    """Patch extraction and the 4-connected patch-grid graph (patch_graph_construction.py:5-136)."""

    def __init__(self, patch_size=16):
        self.patch_size = patch_size

    def grid_dims(self, H: int, W: int) -> Tuple[int, int]:
        p = self.patch_size
        return (H + p - 1) // p, (W + p - 1) // p

    def image_to_patches(self, image_tensor_chw):
        """``(C,H,W) -> (N,C,P,P), (nph,npw)`` (:15-47).  A strided copy; stock torch view ops,
        only needed by callers that want the raw patches (the block itself pools with
        ``get_patch_features``)."""
        C, H, W = image_tensor_chw.shape
        p = self.patch_size
        if H % p != 0 or W % p != 0:
            image_tensor_chw = F.pad(image_tensor_chw, (0, (p - W % p) % p, 0, (p - H % p) % p))
        patches = image_tensor_chw.unfold(1, p, p).unfold(2, p, p)
        nph, npw = patches.shape[1], patches.shape[2]
        patches = patches.permute(1, 2, 0, 3, 4).contiguous().view(-1, C, p, p)
        return patches, (nph, npw)

    def construct_patch_graph(self, image_tensor_chw, patch_features_flat):
        """Returns ``(patch_features_flat, edge_index (2,E) int64)`` (:49-102); the feature tensor is
        returned as the same object, the edge list is generated on the GPU in the reference's order."""
        _, H, W = image_tensor_chw.shape
        nph, npw = self.grid_dims(H, W)
        n = nph * npw
        if patch_features_flat.shape[0] != n:
            raise ValueError(f"Number of patch features ({patch_features_flat.shape[0]}) "
                             f"does not match expected number of patches ({n}) "
                             f"for image {H}x{W} and patch size {self.patch_size}.")
        dev = patch_features_flat.device
        if dev.type != "cuda":
            raise RuntimeError("mingraph_unet_b200 runs on CUDA tensors only (there is no CPU fallback)")
        edge_index = ops.grid_edge_index(nph, npw, dev)
        if edge_index.shape[1] > 0:
            register(edge_index, Graph.grid(nph, npw, dev))
        return patch_features_flat, edge_index

    def get_patch_features(self, feature_map: torch.Tensor, patch: Optional[Tuple[int, int]] = None,
                           out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """Per-channel patch mean ``(C,Hf,Wf) -> (N,C)`` or ``(B,C,Hf,Wf) -> (B,N,C)``: the pooling
        ``get_patch_features_from_unet_encoder`` documents (:104-109, unimplemented there) as
        ``image_to_patches(fm)[0].mean((2,3))``."""
        ph, pw = patch if patch is not None else (self.patch_size, self.patch_size)
        single = feature_map.dim() == 3
        out = ops.pool_patches(feature_map.unsqueeze(0) if single else feature_map, ph, pw, out_dtype)
        return out[0] if single else out

    def get_patch_features_from_unet_encoder(self, unet_encoder_features, patches_coords_info):
        """The reference raises NotImplementedError here (:104-136); same contract.  Use
        :meth:`get_patch_features`."""
        raise NotImplementedError("Feature extraction from U-Net for specific patches is complex and "
                                  "depends on U-Net architecture. Use get_patch_features(feature_map).")


class MinCutRefinement(nn.Module):
    """Soft normalized-cut loss + soft segment assignment (mincut_refinement.py:5-205)."""

    def __init__(self, gamma_unet_priors=0.5, sigma_intensity=10.0, sigma_features=1.0):
        super().__init__()
        self.gamma_unet_priors = gamma_unet_priors
        self.sigma_intensity = sigma_intensity
        self.sigma_features = sigma_features

    def compute_edge_weights_for_ncut(self, node_features, edge_index):
        """``exp(-|f_src - f_tgt|^2 / 2)`` per edge (:30-52; sigma hard-coded to 1.0 at :50); differentiable with
        respect to ``node_features`` like the reference's."""
        return edge_weights_apply(node_features.float(), edge_index)

    def normalized_cut_loss(self, node_features, edge_index, segment_assignments_soft, num_segments_k):
        """(:55-160).  Returns a 0-dim tensor; where the reference returns the Python float ``0.0``
        (no segment with association > 1e-8) this returns a tensor holding 0.0, without a device
        sync."""
        N = node_features.size(0)
        if tuple(segment_assignments_soft.shape) != (N, num_segments_k):
            raise ValueError("segment_assignments_soft shape mismatch.")
        g = edge_index if isinstance(edge_index, Graph) else Graph.from_edge_index(edge_index, N)
        return ncut_loss_apply(node_features.float(), segment_assignments_soft.float(), g).sum()

    def forward(self, gat_refined_patch_features, patch_graph_edge_index, num_expected_segments,
                segment_predictor_network=None):
        if segment_predictor_network is None:
            raise ValueError("segment_predictor_network is required to get segment assignments for Ncut loss.")
        logits = segment_predictor_network(gat_refined_patch_features, patch_graph_edge_index)
        S = softmax_rows(logits.float())
        loss = self.normalized_cut_loss(gat_refined_patch_features, patch_graph_edge_index, S, num_expected_segments)
        return loss, S


class PatchSegmentPredictor(nn.Module):
    """Segment-logit predictor (train_end_to_end.py:40-70): a 1-layer ``GATNetwork`` when
    ``use_gnn`` else a 2-layer MLP (stock torch, not on the hot path)."""

    def __init__(self, in_dim, num_segments, hidden_dim=None, use_gnn=False, num_gnn_layers=1, num_heads=1):
        super().__init__()
        self.use_gnn = use_gnn
        if use_gnn:
            self.gnn_predictor = GATNetwork(node_feature_dim=in_dim, hidden_dim=hidden_dim if hidden_dim else in_dim,
                                            output_dim=num_segments, num_heads=num_heads,
                                            num_gat_layers=num_gnn_layers, dropout_rate=0.1, alpha=0.2)
        else:
            if hidden_dim is None:
                hidden_dim = in_dim * 2
            self.mlp_predictor = nn.Sequential(nn.Linear(in_dim, hidden_dim), nn.ReLU(),
                                               nn.Linear(hidden_dim, num_segments))

    def forward(self, x, edge_index=None):
        if self.use_gnn:
            if edge_index is None:
                raise ValueError("edge_index must be provided for GNN-based segment predictor.")
            return self.gnn_predictor(x, edge_index)
        return self.mlp_predictor(x)


class FeatureConsistencyLoss(nn.Module):
    """``L_feature`` (model/unet/feature_loss.py:5-123): per patch ``y*d^2 + (1-y)*max(0, margin-d)^2``
    with ``d = |f_unet - f_graph|``, summed over patches, averaged over the batch.  Same signature,
    same ``ValueError``s (:95-101); one streaming kernel + a fixed-order reduction, differentiable
    w.r.t. both feature tensors."""

    def __init__(self, margin=1.0):
        super().__init__()
        self.margin = margin

    def forward(self, f_unet, f_graph, correspondence_map_y, regions_unet=None, regions_graph=None):
        B, N_patches, _ = f_unet.shape
        _, _, _ = f_graph.shape
        if f_unet.shape != f_graph.shape:
            raise ValueError(f"f_unet ({f_unet.shape}) and f_graph ({f_graph.shape}) must have same dimensions "
                             f"for this loss version.")
        if correspondence_map_y.shape != (B, N_patches):
            raise ValueError(f"correspondence_map_y (patch_region_labels_y) shape ({correspondence_map_y.shape}) "
                             f"is not (Batch, Num_Patches) = ({B}, {N_patches}).")
        return feature_loss_apply(f_unet, f_graph, correspondence_map_y, float(self.margin))


class TVLoss(nn.Module):
    """Total-variation smoothness loss (scripts/train_end_to_end.py:73-89) over ``(B,C,H,W)``; the map
    is read once by a strip-walking kernel, differentiable w.r.t. ``x``."""

    def __init__(self, weight=1.0):
        super().__init__()
        self.weight = weight

    def forward(self, x):
        return tv_loss_apply(x, float(self.weight))


class FeatureFusion(nn.Module):
    """``F_f = Concat(F_u, F_g)`` (model/fusion_detection/feature_fusion.py:5-153) — the consumer of the graph block's
    output, same constructor and ``forward`` signature (``out`` is an addition).

    * ``f_u_list``: U-Net feature maps; scales whose size differs from ``target_spatial_size`` are resized with
      stock ``F.interpolate(mode='bilinear', align_corners=False)`` (:67-73; the conv side stays in PyTorch).
    * ``f_g`` 4-D ``(B, D, H, W)`` (:134-138, what train_end_to_end.py:439-443 passes): resized the same way if needed.
    * ``f_g`` 2-D ``(R, D)`` + ``region_to_pixel_map (B, H, W)`` (:81-132): every pixel takes the embedding of its region,
      pixels whose index is outside ``[0, R)`` stay zero — one gather kernel (``mg_region_map_gather``) instead of the
      reference's per-image mask / index / scatter sequence.

    The fused tensor is allocated once and every input is written straight into its channel slice (the reference
    materialises ``f_u_combined``, ``f_g_pixel`` and then ``cat`` re-copies both).  An input that already IS the matching
    slice of ``out`` (e.g. the block's un-pool wrote ``F_g`` into ``out[:, C_u:]``, scope row f1) is not copied at all.
    The per-region branch is inference-only for now (no gradient to ``f_g``); the 4-D branch is differentiable through
    torch's slice copies.  CUDA tensors only."""

    def __init__(self, unet_feature_dims, gat_feature_dim, fusion_method="concat"):
        super().__init__()
        self.unet_feature_dims = unet_feature_dims
        self.gat_feature_dim = gat_feature_dim
        self.fusion_method = fusion_method.lower()

    @staticmethod
    def _place(dst: torch.Tensor, src: torch.Tensor, size) -> None:
        if (src.size(2), src.size(3)) != tuple(size):
            src = F.interpolate(src, size=tuple(size), mode="bilinear", align_corners=False)
        if src.data_ptr() == dst.data_ptr() and src.shape == dst.shape and src.stride() == dst.stride() and src.dtype == dst.dtype:
            return                                   # already in place (written there by its producer)
        dst.copy_(src)

    def forward(self, f_u_list, f_g, target_spatial_size=None, region_to_pixel_map=None, out=None):
        B = f_u_list[0].size(0)
        if target_spatial_size is None:
            target_spatial_size = (f_u_list[0].size(2), f_u_list[0].size(3))
        Hh, Ww = int(target_spatial_size[0]), int(target_spatial_size[1])
        per_region = f_g.ndim == 2 and region_to_pixel_map is not None
        if not per_region and f_g.ndim != 4:
            raise ValueError(f"f_g has unsupported shape {f_g.shape}. "
                             "Expected (Num_regions, D_gat) with region_map or (B, D_gat, H, W).")
        if self.fusion_method not in ("concat", "add"):
            raise NotImplementedError(f"Fusion method '{self.fusion_method}' not implemented.")
        c_u = sum(int(t.size(1)) for t in f_u_list)
        d_g = int(self.gat_feature_dim) if per_region else int(f_g.size(1))
        if self.fusion_method == "add" and c_u != d_g:
            raise ValueError("Channel dimensions must match for 'add' fusion or implement adaptation.")
        ops._need_cuda(*f_u_list, f_g, region_to_pixel_map if per_region else None)     # no CPU fallback
        dev = f_u_list[0].device
        dt = f_u_list[0].dtype
        for t in f_u_list[1:]:
            dt = torch.promote_types(dt, t.dtype)
        dt = torch.promote_types(dt, torch.float32 if per_region else f_g.dtype)     # f_g_pixel is float32 zeros (:85)
        if per_region:
            if f_g.requires_grad and torch.is_grad_enabled():
                raise NotImplementedError("FeatureFusion: the per-region branch is inference-only (pass the 4-D f_g, as "
                                          "scripts/train_end_to_end.py does, to train through the fusion)")
            if f_g.size(1) != d_g:
                raise RuntimeError(f"shape mismatch: f_g has {f_g.size(1)} columns, gat_feature_dim is {d_g}")
            rmap = region_to_pixel_map
            if rmap.dim() != 3 or rmap.size(0) != B or (rmap.size(1), rmap.size(2)) != (Hh, Ww):
                raise IndexError(f"region_to_pixel_map {tuple(rmap.shape)} does not match (B, H_out, W_out) = "
                                 f"({B}, {Hh}, {Ww})")
            if rmap.dtype not in (torch.int32, torch.int64):
                rmap = rmap.long()                                                   # (:115)
            table = f_g.float()
        if self.fusion_method == "add":
            f_u = f_u_list[0] if len(f_u_list) == 1 and (f_u_list[0].size(2), f_u_list[0].size(3)) == (Hh, Ww) else None
            if f_u is None:
                f_u = torch.empty((B, c_u, Hh, Ww), dtype=dt, device=dev)
                c = 0
                for t in f_u_list:
                    self._place(f_u[:, c:c + t.size(1)], t, (Hh, Ww))
                    c += t.size(1)
            if per_region:
                g = ops.region_map_gather(table, rmap, out_dtype=torch.float32)
            else:
                g = f_g if (f_g.size(2), f_g.size(3)) == (Hh, Ww) else \
                    F.interpolate(f_g, size=(Hh, Ww), mode="bilinear", align_corners=False)
            return f_u + g
        if out is None:
            out = torch.empty((B, c_u + d_g, Hh, Ww), dtype=dt, device=dev)
        elif tuple(out.shape) != (B, c_u + d_g, Hh, Ww) or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous ({B}, {c_u + d_g}, {Hh}, {Ww}) tensor")
        c = 0
        for t in f_u_list:
            self._place(out[:, c:c + t.size(1)], t, (Hh, Ww))
            c += t.size(1)
        if per_region:
            ops.region_map_gather(table, rmap, out=out[:, c_u:])
        else:
            self._place(out[:, c_u:], f_g, (Hh, Ww))
        return out
