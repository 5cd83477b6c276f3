"""``RefGraphBlock``: the reference's per-image graph loop (scripts/train_end_to_end.py:300-425) driven through the
UNTOUCHED reference classes (TEST INFRASTRUCTURE — the CPU arm ``bench.py --impl reference`` times, and a second
checker beside ``oracle/restate.py``).  Nothing of the reference is re-implemented here: every stage is a call into
``PatchGraphConstructor``, ``GATNetwork``, ``PatchSegmentPredictor`` and ``MinCutRefinement`` as imported by
``oracle/ref_loader.py``, in the script's own order, with the script's own glue expressions (cited line by line).
Differences from the script, both forced: node features are passed in / pooled from a feature map instead of the
``torch.randn`` placeholder (:326), and the feature-loss call (:344) is omitted because it raises in the reference
(SURVEY 0.4)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import ref_loader


class RefGraphBlock:
    def __init__(self, node_feature_dim: int = 20, gat_hidden_dim: int = 128, gat_output_dim: int = 64, num_heads: int = 4,
                 num_segments: int = 2, patch_size: int = 16, seed: Optional[int] = 1234):
        R = ref_loader.load()
        if seed is not None:
            torch.manual_seed(seed)
        self.K, self.D, self.patch_size = num_segments, gat_output_dim, patch_size
        self.patch_graph_constructor = R.PatchGraphConstructor(patch_size)                                  # :118
        self.patch_gat_model = R.GATNetwork(node_feature_dim, gat_hidden_dim, gat_output_dim, num_heads, 1, 0.1, 0.2).eval()  # :144-152
        if R.PatchSegmentPredictor is not None:                                                             # :156-163
            self.segment_predictor = R.PatchSegmentPredictor(gat_output_dim, num_segments, gat_output_dim // 2, use_gnn=True,
                                                             num_gnn_layers=1, num_heads=max(1, num_heads // 2)).eval()
            self.predictor_net = self.segment_predictor.gnn_predictor
        else:       # scripts/ not importable (cv2 / tqdm missing): the use_gnn=True branch is exactly this network (:43-54)
            self.segment_predictor = self.predictor_net = R.GATNetwork(gat_output_dim, gat_output_dim // 2, num_segments,
                                                                       max(1, num_heads // 2), 1, 0.1, 0.2).eval()
        self.mincut_module = R.MinCutRefinement().eval()                                                    # :164
        self.region_gat_model = R.GATNetwork(gat_output_dim, gat_hidden_dim, gat_output_dim, num_heads, 1, 0.1, 0.2).eval()  # :170-178

    def nets(self):
        return {"patch": self.patch_gat_model, "pred": self.predictor_net, "region": self.region_gat_model}

    @torch.no_grad()
    def pooled_node_features(self, feature_map_chw: torch.Tensor) -> torch.Tensor:
        """Patch-mean pooling with the reference's own patch extractor (patch_graph_construction.py:15-47) and the
        mean idiom of scripts/graph_refinement.py:78,98 kept per channel (docstring intent of :104-109)."""
        patches, _ = self.patch_graph_constructor.image_to_patches(feature_map_chw)
        return patches.mean(dim=(2, 3))

    @torch.no_grad()
    def image(self, H: int, W: int, node_features: Optional[torch.Tensor] = None,
              feature_map: Optional[torch.Tensor] = None, want_dense: bool = True) -> Dict[str, torch.Tensor]:
        """One iteration of the loop at :300 for an (H, W) image."""
        K, D = self.K, self.D
        image_for_dims = torch.zeros(1, H, W) if feature_map is None else feature_map
        _, (num_patches_h, num_patches_w) = self.patch_graph_constructor.image_to_patches(image_for_dims)   # :318-319
        x = node_features if node_features is not None else self.pooled_node_features(feature_map)          # (:326)
        _, edge_index = self.patch_graph_constructor.construct_patch_graph(image_for_dims, x)               # :329
        f_g_patches = self.patch_gat_model(x, edge_index)                                                   # :332
        l_partition, soft_assign = self.mincut_module(f_g_patches, edge_index, K, self.segment_predictor)   # :348
        hard_assign = torch.argmax(soft_assign, dim=1)                                                      # :356
        region_features = torch.zeros(K, D)                                                                 # :367
        for k in range(K):                                                                                  # :368-373
            mask = hard_assign == k
            if mask.sum() > 0:
                region_features[k] = f_g_patches[mask].mean(dim=0)
        s, t = torch.triu_indices(K, K, offset=1)                                                           # :376-378
        region_edge_index = torch.stack([torch.cat([s, t]), torch.cat([t, s])], dim=0)
        if region_edge_index.shape[1] > 0:                                                                  # :383-389
            f_g_region = self.region_gat_model(region_features, region_edge_index)
        else:
            f_g_region = region_features
        out = {"h": f_g_patches, "S": soft_assign, "hard": hard_assign, "R": region_features, "G": f_g_region,
               "loss": torch.as_tensor(float(l_partition), dtype=torch.float32), "edge_index": edge_index,
               "grid": (num_patches_h, num_patches_w)}
        if want_dense:
            f_g_patch_level = f_g_region[hard_assign]                                                       # :403-406
            f_g_map = f_g_patch_level.T.reshape(D, num_patches_h, num_patches_w)                            # :411
            out["f_g"] = F.interpolate(f_g_map.unsqueeze(0), size=(H, W), mode="nearest").squeeze(0)        # :417-421
        return out
