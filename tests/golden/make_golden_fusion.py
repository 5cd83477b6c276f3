"""Golden fixtures for ``FeatureFusion`` (the consumer of the graph block's output, scope row f1), generated from
the UNTOUCHED reference class ``model/fusion_detection/feature_fusion.py``.  Run in the build container only
(``/root/reference`` is mounted there):

    python tests/golden/make_golden_fusion.py

Cases: the per-region branch (:81-132) with int64 / int32 maps, invalid indices (negative and >= R: those pixels stay
zero), odd sizes that force the scalar kernel, a superpixel-like map with constant runs; the 4-D branch at equal and at
different resolution (bilinear resize, :134-138); two U-Net scales (:67-75); ``add`` fusion (:144-148).
"""
from __future__ import annotations

import importlib
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_loader  # noqa: E402


def main():
    ref_loader.load()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ff = importlib.import_module("model.fusion_detection.feature_fusion")
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(606)
    d = {}

    def blocky_map(B, H, W, R, cell):
        """superpixel-like: constant cells of ``cell`` x ``cell`` pixels with random region ids"""
        ids = torch.randint(0, R, (B, -(-H // cell), -(-W // cell)), generator=g)
        return ids.repeat_interleave(cell, 1).repeat_interleave(cell, 2)[:, :H, :W].contiguous()

    region_cases = {
        # tag: (B, Cu, H, W, R, D, map kind, map dtype)
        "rand_i64": (2, 5, 16, 32, 7, 64, "random", torch.int64),
        "rand_i32": (2, 3, 16, 24, 9, 64, "random", torch.int32),
        "invalid": (2, 4, 8, 16, 5, 32, "invalid", torch.int64),        # -1 and >= R entries
        "blocky": (3, 8, 32, 32, 6, 64, "blocky", torch.int64),
        "odd": (1, 2, 7, 13, 4, 6, "random", torch.int64),              # W % 4 != 0, D % 4 != 0
        "allbad": (1, 2, 4, 8, 3, 8, "allbad", torch.int64),            # no valid pixel: F_g part all zero
    }
    for tag, (B, Cu, H, W, R, D, kind, mdt) in region_cases.items():
        fu = torch.randn(B, Cu, H, W, generator=g)
        table = torch.randn(R, D, generator=g)
        if kind == "blocky":
            m = blocky_map(B, H, W, R, 8)
        elif kind == "allbad":
            m = torch.full((B, H, W), R + 2, dtype=torch.int64)
        else:
            m = torch.randint(0, R, (B, H, W), generator=g)
            if kind == "invalid":
                bad = torch.rand(B, H, W, generator=g)
                m = torch.where(bad < 0.2, torch.full_like(m, -1), m)
                m = torch.where(bad > 0.85, torch.full_like(m, R + 3), m)
        m = m.to(mdt)
        out = ff.FeatureFusion([Cu], D, "concat")([fu], table, region_to_pixel_map=m)
        d[f"rg_{tag}_fu"], d[f"rg_{tag}_table"], d[f"rg_{tag}_map"] = fu.numpy(), table.numpy(), m.numpy()
        d[f"rg_{tag}_out"] = out.numpy()

    # 4-D branch, same size (what scripts/train_end_to_end.py:439-443 passes)
    fu, fg = torch.randn(2, 6, 16, 16, generator=g), torch.randn(2, 10, 16, 16, generator=g)
    d["d4_same_fu"], d["d4_same_fg"] = fu.numpy(), fg.numpy()
    d["d4_same_out"] = ff.FeatureFusion([6], 10)([fu], fg, target_spatial_size=(16, 16)).numpy()
    # 4-D branch, F_g at half resolution, two U-Net scales (one at quarter resolution)
    fu0, fu1 = torch.randn(2, 4, 16, 24, generator=g), torch.randn(2, 3, 4, 6, generator=g)
    fg = torch.randn(2, 8, 8, 12, generator=g)
    d["d4_resize_fu0"], d["d4_resize_fu1"], d["d4_resize_fg"] = fu0.numpy(), fu1.numpy(), fg.numpy()
    d["d4_resize_out"] = ff.FeatureFusion([4, 3], 8)([fu0, fu1], fg).numpy()
    # add fusion, per-region branch
    fu, table = torch.randn(2, 16, 8, 8, generator=g), torch.randn(5, 16, generator=g)
    m = torch.randint(-1, 6, (2, 8, 8), generator=g)
    d["add_fu"], d["add_table"], d["add_map"] = fu.numpy(), table.numpy(), m.numpy()
    d["add_out"] = ff.FeatureFusion([16], 16, "add")([fu], table, region_to_pixel_map=m).numpy()

    path = os.path.join(HERE, "fusion.npz")
    np.savez_compressed(path, **d)
    print("wrote", path, os.path.getsize(path), "bytes,", len(d), "arrays")


if __name__ == "__main__":
    main()
