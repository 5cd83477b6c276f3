"""Batched graph block: the per-image loop of scripts/train_end_to_end.py:300-425 for all ``B``
images at once (block-diagonal graphs, O(10) kernel launches per batch, no host syncs).

Stage order (reference lines in parentheses):
  node features  — patch-mean pool of a feature map (patch_graph_construction.py:104-109 intent;
                   graph_refinement.py:78,98,103 idiom) or given ``(B,N,in)`` (:326 placeholder)
  patch graph    — 4-connected grid (:329)
  patch GAT      — ``patch_gat_model`` (:332)
  min-cut        — predictor GAT -> softmax -> N-cut loss (:348; mincut_refinement.py:192-196)
  hard labels    — argmax (:356)
  region pool    — per-label mean, zeros for empty regions (:368-373)
  region GAT     — complete digraph over K regions (:376-389)
  un-pool        — gather by label, nearest up-sampling to (D,H,W) (:403-421), stacked (B,D,H,W) (:432)
"""
from __future__ import annotations

from typing import NamedTuple, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .autograd import feature_loss_apply, ncut_loss_apply, segment_mean_apply, softmax_rows_with_labels, unpool_apply
from .graph import Graph
from .modules import GATNetwork, MinCutRefinement, PatchGraphConstructor, PatchSegmentPredictor, _layer_forward


class GraphBlockOutput(NamedTuple):
    f_g: Optional[torch.Tensor]          # (B, D, H, W) dense graph features (train_end_to_end.py:432)
    l_partition: torch.Tensor            # (B,) per-image N-cut loss (the reference averages them, :428)
    soft_assignments: torch.Tensor       # (B, N, K)
    hard_labels: torch.Tensor            # (B, N) int32
    patch_features: torch.Tensor         # (B, N, D) patch-GAT output h
    region_features: torch.Tensor        # (B, K, D) region-GAT output
    grid: Tuple[int, int]                # (nph, npw)
    l_feature: Optional[torch.Tensor] = None   # 0-dim: FeatureConsistencyLoss(f_unet_patches, patch_features, y) (:344)


class GraphBlock(nn.Module):
    """Sub-module names follow the reference's variables (train_end_to_end.py:144-178) so their
    state_dicts map one to one: ``patch_gat_model``, ``segment_predictor``, ``mincut_module``,
    ``region_gat_model``."""

    def __init__(self, node_feature_dim: int = 20, gat_hidden_dim: int = 128, gat_output_dim: int = 64,
                 num_heads: int = 4, num_segments: int = 2, patch_size: int = 16, dropout_rate: float = 0.1,
                 alpha: float = 0.2):
        super().__init__()
        self.patch_size = patch_size
        self.num_segments = num_segments
        self.gat_output_dim = gat_output_dim
        self.patch_graph_constructor = PatchGraphConstructor(patch_size)
        self.patch_gat_model = GATNetwork(node_feature_dim, gat_hidden_dim, gat_output_dim, num_heads, 1,
                                          dropout_rate, alpha)
        self.segment_predictor = PatchSegmentPredictor(gat_output_dim, num_segments, hidden_dim=gat_output_dim // 2,
                                                       use_gnn=True, num_gnn_layers=1, num_heads=max(1, num_heads // 2))
        self.mincut_module = MinCutRefinement()
        self.region_gat_model = GATNetwork(gat_output_dim, gat_hidden_dim, gat_output_dim, num_heads, 1,
                                           dropout_rate, alpha)
        self.fused = True            # False: compose the stand-alone kernels (same results, more launches)
        self._prep_cache = None

    def forward(self, node_features: Optional[torch.Tensor] = None, image_size: Optional[Tuple[int, int]] = None,
                feature_map: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                out_dtype: Optional[torch.dtype] = None, want_dense: bool = True, _block_outs=None,
                f_unet_patches: Optional[torch.Tensor] = None, patch_labels_y: Optional[torch.Tensor] = None,
                feature_loss_margin: float = 1.0, _after_block=None, _peer=None) -> GraphBlockOutput:
        """Either ``node_features (B,N,in)`` + ``image_size (H,W)`` or a per-pixel ``feature_map
        (B,in,H,W)`` (patch-mean pooled to node features).  ``out`` may be a channel slice of a fusion
        buffer ``(B,Ctot,H,W)[:, c0:c0+D]``; the dense map is written there directly.
        ``f_unet_patches (B,N,D)`` + ``patch_labels_y (B,N)``: also return ``l_feature``, the reference's
        feature-consistency loss between them and the patch-GAT output (train_end_to_end.py:344,
        batch mean of the per-image sums).  ``_peer``: a ``distributed.PeerExchange.slot(i)`` descriptor — the block
        kernel then also pushes the small outputs to every rank (multi-GPU exchange fused into the launch)."""
        if f_unet_patches is not None and patch_labels_y is None:
            raise ValueError("patch_labels_y is required with f_unet_patches")
        fl_in = None
        if f_unet_patches is not None and f_unet_patches.dtype == torch.float32 and f_unet_patches.is_cuda \
                and not (torch.is_grad_enabled() and f_unet_patches.requires_grad):
            fl_in = (f_unet_patches, patch_labels_y, float(feature_loss_margin))     # folded into the block kernel when it runs
        res = self._forward(node_features, image_size, feature_map, out, out_dtype, want_dense, _block_outs, _after_block,
                            _peer, fl_in)
        if f_unet_patches is None or res.l_feature is not None:
            return res
        if tuple(f_unet_patches.shape) != tuple(res.patch_features.shape):
            raise ValueError(f"f_unet ({f_unet_patches.shape}) and f_graph ({res.patch_features.shape}) must have "
                             f"same dimensions for this loss version.")
        return res._replace(l_feature=feature_loss_apply(f_unet_patches, res.patch_features, patch_labels_y,
                                                         float(feature_loss_margin)))

    def _forward(self, node_features, image_size, feature_map, out, out_dtype, want_dense, _block_outs,
                 _after_block=None, _peer=None, _feature_loss=None) -> GraphBlockOutput:
        if (node_features is None) == (feature_map is None):
            raise ValueError("pass exactly one of node_features / feature_map")
        if feature_map is not None:
            if image_size is None:
                image_size = tuple(feature_map.shape[-2:])
            fh, fw = feature_map.shape[-2:]
            H, W = image_size
            nph, npw = self.patch_graph_constructor.grid_dims(H, W)
            # pooling window in feature-map pixels (patch_size / stride of the feature map)
            if H % fh or W % fw or self.patch_size % (H // fh) or self.patch_size % (W // fw):
                raise ValueError("feature_map resolution must divide the image size and the patch size")
            ph, pw = self.patch_size // (H // fh), self.patch_size // (W // fw)
            node_features = ops.pool_patches(feature_map, ph, pw)
            if node_features.shape[1] != nph * npw:
                raise ValueError("pooled patch count does not match the patch grid")
        if image_size is None:
            raise ValueError("image_size=(H, W) is required with node_features")
        H, W = image_size
        if node_features.dim() != 3:
            raise ValueError("node_features must be (B, N, in)")
        B, N, _ = node_features.shape
        nph, npw = self.patch_graph_constructor.grid_dims(H, W)
        if N != nph * npw:                                            # patch_graph_construction.py:71-74
            raise ValueError(f"Number of patch features ({N}) does not match expected number of patches "
                             f"({nph * npw}) for image {H}x{W} and patch size {self.patch_size}.")
        dev = node_features.device
        K, D = self.num_segments, self.gat_output_dim
        if nph * npw < 2:
            raise RuntimeError("max(): Expected reduction dim to be specified for input.numel() == 0 "
                               "(a 1x1 patch grid has no edges)")
        dense_dtype = out_dtype if out_dtype is not None else node_features.dtype
        layers = (self.patch_gat_model.gat_layers[0], self.segment_predictor.gnn_predictor.gat_layers[0],
                  self.region_gat_model.gat_layers[0])
        needs_autograd = torch.is_grad_enabled() and (node_features.requires_grad or
                                                      any(p.requires_grad for p in self.parameters()))
        use_fused = (self.fused and not needs_autograd and not (self.training and any(l.dropout_rate > 0 for l in layers))
                     and ops.block_supported(B, nph, npw, node_features.shape[-1], D, layers[0].num_heads,
                                             layers[1].num_heads, layers[2].num_heads, K))
        if _peer is not None and not use_fused:
            raise RuntimeError("the fused peer exchange needs the one-launch block kernel (inference, supported shape); "
                               "use distributed.InlineGather otherwise")
        if use_fused:
            # ONE launch: patch GAT -> predictor GAT -> softmax/argmax -> N-cut -> region pool -> region GAT
            l_feature = None
            if _feature_loss is not None and tuple(_feature_loss[0].shape) != (B, N, D):
                raise ValueError(f"f_unet ({_feature_loss[0].shape}) and f_graph ({(B, N, D)}) must have "
                                 f"same dimensions for this loss version.")
            r = ops.block_forward(
                node_features, nph, npw, self._prepared(), D, layers[0].num_heads, layers[1].num_heads,
                layers[2].num_heads, K, slopes=tuple(l.alpha for l in layers), outs=_block_outs, peer=_peer,
                feature_loss=_feature_loss)
            h, S, labels, loss, _, G = r[:6]
            if _feature_loss is not None:
                l_feature = r[6].mean()                 # batch mean of the per-image sums (feature_loss.py:124)
            if _after_block is not None:
                _after_block()          # the small outputs are final here; the un-pool below only reads them
        elif needs_autograd or (self.training and any(l.dropout_rate > 0 for l in layers)):
            # training: the same stages as differentiable ops (csrc/gat_backward.cu, ncut.cu, block_backward.cu)
            g = Graph.grid(nph, npw, dev, B)
            x = node_features.reshape(B * N, -1)
            h = _train_layer(layers[0], x, g)                                           # :332
            logits = _train_layer(layers[1], h, g)                                      # mincut_refinement.py:192
            S, labels = softmax_rows_with_labels(logits)                                # :193, train_end_to_end.py:356
            loss = ncut_loss_apply(h, S, g)                                             # :196
            R = segment_mean_apply(h.view(B, N, D), labels.view(B, N), K)               # :368-373
            if K > 1:                                                                   # :383-389
                G = _train_layer(layers[2], R.reshape(B * K, D), Graph.complete(K, dev, B)).view(B, K, D)
            else:
                G = R
            h, S, labels = h.view(B, N, D), S.view(B, N, K), labels.view(B, N)
            f_g = unpool_apply(G, labels, nph, npw, H, W, out=out, out_dtype=dense_dtype) if want_dense else None
            return GraphBlockOutput(f_g, loss, S, labels, h, G, (nph, npw))
        else:
            g = Graph.grid(nph, npw, dev, B)
            x = node_features.reshape(B * N, -1)
            h = _batched_layer(layers[0], x, g, torch.float32)                          # :332
            logits = _batched_layer(layers[1], h, g, torch.float32)                     # mincut_refinement.py:192
            S, labels = ops.softmax_argmax(logits)                                      # :193, train_end_to_end.py:356
            loss = ops.ncut_loss(h, S, g.rowptr_out, g.col_out, g.nodes_per_graph)      # :196
            R = ops.segment_mean(h.view(B, N, D), labels.view(B, N), K)                 # :368-373
            if K > 1:                                                                   # :383-389
                rg = Graph.complete(K, dev, B)
                G = _batched_layer(layers[2], R.view(B * K, D), rg, torch.float32).view(B, K, D)
            else:
                G = R
            h, S, labels = h.view(B, N, D), S.view(B, N, K), labels.view(B, N)
        f_g = None
        if want_dense:                                                                  # :403-421
            f_g = ops.unpool_nearest(G, labels, nph, npw, H, W, out=out, out_dtype=dense_dtype)
        return GraphBlockOutput(f_g, loss, S, labels, h, G, (nph, npw), l_feature if use_fused else None)

    def __setattr__(self, name, value):
        if isinstance(value, (nn.Module, nn.Parameter)):
            self.__dict__["_weight_mods"] = None          # a sub-network was replaced: re-walk the module tree
        super().__setattr__(name, value)

    def _weight_linears(self):
        """The 2 x heads ``nn.Linear`` modules of the three 1-layer networks, in blob order.  Walking the module tree
        (``net.parameters()``) costs ~130 us of host time per call — as much as a whole pipelined step — so the walk is
        cached; the parameters themselves are read from the modules' dicts on every call, so in-place updates,
        ``load_state_dict`` (also ``assign=True``) and ``.to()`` are all seen.  Replacing a head module inside a
        network (structural surgery) needs ``block._weight_mods = None``."""
        mods = self.__dict__.get("_weight_mods")
        if mods is None:
            mods = []
            for net in (self.patch_gat_model, self.segment_predictor.gnn_predictor, self.region_gat_model):
                for hd in net.gat_layers[0].heads:
                    mods.append(hd.W)
                    mods.append(hd.a)
            self.__dict__["_weight_mods"] = mods
        return mods

    def _weight_key(self):
        key = []
        for m in self._weight_linears():
            p = m._parameters["weight"]
            key.append((p.data_ptr(), p._version))
        return tuple(key)

    def _prepared(self) -> torch.Tensor:
        """Weights re-arranged for the fused kernel, cached per weight version (one tiny launch when
        any parameter changed, e.g. after an optimizer step or ``load_state_dict``)."""
        key = self._weight_key()
        if self._prep_cache is None or self._prep_cache[0] != key:
            nets = (self.patch_gat_model, self.segment_predictor.gnn_predictor, self.region_gat_model)
            stacks = []
            for net in nets:
                heads = list(net.gat_layers[0].heads)
                stacks.append(torch.stack([hd.W.weight.detach() for hd in heads], 0))
                stacks.append(torch.stack([hd.a.weight.detach().view(-1) for hd in heads], 0))
            self._prep_cache = (key, ops.block_prepare(*stacks, out=None if self._prep_cache is None else self._prep_cache[1]))
        return self._prep_cache[1]


def _train_layer(layer, x: torch.Tensor, g: Graph) -> torch.Tensor:
    """Differentiable multi-head layer on a block-diagonal batch: attention dropout inside the kernel, output
    dropout (graph_attention.py:160) by ``nn.Dropout``; fp32 between stages like the inference path."""
    out = _layer_forward(list(layer.heads), x, g, layer.concat, layer.alpha,
                         att_dropout=layer.dropout_rate if layer.training else 0.0, out_dtype=torch.float32)
    return layer.dropout(out)


def _batched_layer(layer, x: torch.Tensor, g: Graph, out_dtype: torch.dtype) -> torch.Tensor:
    """Eval-mode multi-head layer on a block-diagonal batch (dropout is identity in eval; the
    training path goes through ``modules._layer_forward``)."""
    if layer.training and layer.dropout_rate > 0:
        return layer(x, g)
    heads = list(layer.heads)
    W = torch.stack([hd.W.weight for hd in heads], 0).detach()
    a = torch.stack([hd.a.weight.view(-1) for hd in heads], 0).detach()
    return ops.gat_forward(x, g.rowptr_in, g.col_in, W, a, concat=layer.concat, slope=layer.alpha,
                           nodes_per_graph=g.nodes_per_graph, out_dtype=out_dtype)
