"""Golden fixtures for scope row f4 (FeatureConsistencyLoss, TVLoss, multi-layer GAT stack), generated
from the UNTOUCHED reference.  Run in the build container only (``/root/reference`` is mounted there):

    python tests/golden/make_golden_f4.py

* ``FeatureConsistencyLoss`` is imported from ``model/unet/feature_loss.py`` and called with the
  ``(B,N,D), (B,N,D), (B,N)`` arguments its code accepts (``:95-101``).
* ``TVLoss`` is imported from ``scripts/train_end_to_end.py`` (``:73-89``).
* The multi-layer stack composes reference ``MultiHeadGATLayer`` objects with the widths the
  concatenating layer really emits; the reference's own ``GATNetwork(num_gat_layers>=2)`` crashes at
  forward (``graph_attention.py:176-186``), which the fixture also records.
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
    R = ref_loader.load()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fl = importlib.import_module("model.unet.feature_loss")
        te = importlib.import_module("scripts.train_end_to_end")
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(404)
    d = {}

    # ---- FeatureConsistencyLoss ---------------------------------------------------------
    for tag, (B, N, D, margin, scale) in {
        "b3_n37_d64": (3, 37, 64, 1.0, 0.1),        # distances around the margin: both branches active
        "b1_n256_d64": (1, 256, 64, 1.0, 0.12),     # the call at train_end_to_end.py:344 (one image, 16x16 patches)
        "b2_n50_d7": (2, 50, 7, 2.5, 1.0),          # odd width, other margin
    }.items():
        fu = scale * torch.randn(B, N, D, generator=g)
        fg = scale * torch.randn(B, N, D, generator=g)
        y = torch.randint(0, 2, (B, N), generator=g)
        if tag == "b3_n37_d64":
            fg[0, 0] = fu[0, 0]                      # a zero distance (sqrt epsilon path, y = 0 and y = 1)
            fg[0, 1] = fu[0, 1]
            y[0, 0], y[0, 1] = 0, 1
        fu.requires_grad_(True)
        fg.requires_grad_(True)
        loss = fl.FeatureConsistencyLoss(margin=margin)(fu, fg, y)
        loss.backward()
        d[f"fl_{tag}_fu"], d[f"fl_{tag}_fg"], d[f"fl_{tag}_y"] = fu.detach().numpy(), fg.detach().numpy(), y.numpy()
        d[f"fl_{tag}_margin"] = np.float32(margin)
        d[f"fl_{tag}_loss"] = np.float32(loss.item())
        d[f"fl_{tag}_gfu"], d[f"fl_{tag}_gfg"] = fu.grad.numpy(), fg.grad.numpy()

    # ---- TVLoss -----------------------------------------------------------------------------
    for tag, (B, C, H, W, weight) in {
        "b2_c1_64x64": (2, 1, 64, 64, 1.0),         # a (B,1,H,W) probability map, as at train_end_to_end.py:461
        "b3_c2_37x53": (3, 2, 37, 53, 0.5),         # ragged: scalar path, partial strips
        "b1_c3_5x200": (1, 3, 5, 200, 2.0),         # wide and short: several column chunks
        "b2_c2_19x8": (2, 2, 19, 8, 1.0),
    }.items():
        x = torch.rand(B, C, H, W, generator=g).requires_grad_(True)
        loss = te.TVLoss(weight)(x)
        loss.backward()
        d[f"tv_{tag}_x"] = x.detach().numpy()
        d[f"tv_{tag}_weight"] = np.float32(weight)
        d[f"tv_{tag}_loss"] = np.float32(loss.item())
        d[f"tv_{tag}_gx"] = x.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "losses.npz"), **d)

    # ---- multi-layer GAT stack ------------------------------------------------------------------
    m = {}
    pgc = R.PatchGraphConstructor(16)
    for tag, (hp, wp, fin, hidden, fout, heads, nl) in {
        "2layer": (6, 7, 20, 32, 64, 4, 2),
        "3layer": (5, 5, 12, 16, 10, 2, 3),
    }.items():
        torch.manual_seed(77 + nl)
        widths = [fin] + [hidden] * (nl - 1)
        layers = [R.MultiHeadGATLayer(widths[i], hidden, heads, 0.1, 0.2, concat=True).eval() for i in range(nl - 1)]
        layers.append(R.MultiHeadGATLayer(hidden, fout, heads, 0.1, 0.2, concat=False).eval())
        x = torch.randn(hp * wp, fin, generator=g)
        _, ei = pgc.construct_patch_graph(torch.zeros(1, hp * 16, wp * 16), x)
        with torch.no_grad():
            h = x
            for l in layers:
                h = l(h, ei)
        m[f"{tag}_x"], m[f"{tag}_y"] = x.numpy(), h.numpy()
        m[f"{tag}_meta"] = np.array([hp, wp, fin, hidden, fout, heads, nl])
        for i, l in enumerate(layers):
            sd = l.state_dict()
            m[f"{tag}_W{i}"] = np.stack([sd[f"heads.{k}.W.weight"].numpy() for k in range(heads)])
            m[f"{tag}_a{i}"] = np.stack([sd[f"heads.{k}.a.weight"].numpy().reshape(-1) for k in range(heads)])
        # the reference's own >=2-layer network fails at forward
        try:
            with torch.no_grad():
                R.GATNetwork(fin, hidden, fout, heads, nl, 0.1, 0.2).eval()(x, ei)
            m[f"{tag}_ref_network_error"] = np.array("")
        except RuntimeError as e:
            m[f"{tag}_ref_network_error"] = np.array(str(e)[:120])
    np.savez_compressed(os.path.join(HERE, "multilayer_gat.npz"), **m)
    for f in ("losses.npz", "multilayer_gat.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))
    print({k: str(v) for k, v in m.items() if k.endswith("error")})


if __name__ == "__main__":
    main()
