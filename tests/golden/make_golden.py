"""Generate the golden fixtures in this directory from the UNTOUCHED reference.

Run in the build container only (``/root/reference`` is mounted there):

    python tests/golden/make_golden.py

It imports the reference classes in place (``oracle/ref_loader.py``), runs them on
seeded CPU fp32 inputs in ``eval()`` mode and stores inputs, weights and outputs
as ``.npz``.  The per-image block case composes the reference classes exactly in
the stage order of ``scripts/train_end_to_end.py:318-421`` (node features passed
in instead of the ``randn`` placeholder at ``:326``; the feature-loss call at
``:344`` crashes in the reference and is omitted).
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_loader  # noqa: E402


def sd_stacks(net):
    sd = net.state_dict()
    H = len(net.gat_layers[0].heads)
    W = np.stack([sd[f"gat_layers.0.heads.{h}.W.weight"].numpy() for h in range(H)])
    a = np.stack([sd[f"gat_layers.0.heads.{h}.a.weight"].numpy().reshape(-1) for h in range(H)])
    return W, a


def sha16(arr: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()[:16]


def main():
    R = ref_loader.load()
    torch.set_num_threads(1)
    out = {}

    # ---- KAT-1: analytic single head -------------------------
    head = R.GraphAttentionLayer(2, 2, 0.0, 0.2).eval()
    with torch.no_grad():
        head.W.weight.copy_(torch.eye(2))
        head.a.weight.copy_(torch.tensor([[1.0, 0.0, 0.0, 1.0]]))
        x = torch.tensor([[1.0, 0.0], [0.0, 2.0], [-1.0, -1.0]])
        ei = torch.tensor([[0, 1, 2, 0], [1, 0, 0, 2]])
        y = head(x, ei)
    np.savez(os.path.join(HERE, "kat1_head.npz"), x=x.numpy(), ei=ei.numpy(), y=y.numpy())

    # ---- KAT-2: default patch GAT on a 16x16 grid ---------------------------------
    torch.manual_seed(1234)
    net = R.GATNetwork(20, 128, 64, 4, 1, 0.1, 0.2).eval()
    x = torch.randn(256, 20)
    pgc = R.PatchGraphConstructor(16)
    _, ei = pgc.construct_patch_graph(torch.zeros(3, 256, 256), x)
    with torch.no_grad():
        y = net(x, ei)
    W, a = sd_stacks(net)
    np.savez(os.path.join(HERE, "kat2_patch_gat.npz"), x=x.numpy(), W=W, a=a, y=y.numpy(),
             ei_sha=np.array(sha16(ei.numpy())))

    # ---- KAT-3: edge_index for many grid shapes -----------------------------------
    shapes = [(1, 1), (1, 4), (4, 1), (2, 2), (3, 5), (4, 4), (5, 5), (7, 7), (16, 16), (32, 32), (64, 64), (9, 13)]
    d = {}
    for hp, wp in shapes:
        img = torch.zeros(1, hp * 16 - 3 if hp > 1 else 16, wp * 16 - 5 if wp > 1 else 16)   # non-divisible sizes
        _, e = pgc.construct_patch_graph(img, torch.zeros(hp * wp, 1))
        assert e.dtype == torch.int64
        d[f"sha_{hp}x{wp}"] = np.array(sha16(e.numpy()))
        d[f"shape_{hp}x{wp}"] = np.array(e.shape)
        if hp * wp <= 64:
            d[f"ei_{hp}x{wp}"] = e.numpy()
    np.savez(os.path.join(HERE, "kat3_edge_index.npz"), **d)

    # ---- multi-head layers: concat / average, odd dims, random multigraph -------------
    g = torch.Generator().manual_seed(7)
    N, E = 53, 311
    ei = torch.randint(0, N - 3, (2, E), generator=g)            # nodes N-3.. have no edges at all
    ei[1, ei[1] == 5] = 6                                        # node 5 has zero in-degree
    ei[:, 10] = ei[:, 11]                                        # a duplicated edge
    ei[0, 20] = ei[1, 20]                                        # a self loop
    cases = {}
    for name, (fin, fout, heads, concat) in {
        "avg_17_24_3": (17, 24, 3, False),
        "cat_33_32_4": (33, 32, 4, True),
        "avg_64_2_2": (64, 2, 2, False),
        "avg_130_40_1": (130, 40, 1, False),
    }.items():
        torch.manual_seed(100 + fin)
        layer = R.MultiHeadGATLayer(fin, fout, heads, 0.1, 0.2, concat=concat).eval()
        x = torch.randn(N, fin, generator=g)
        with torch.no_grad():
            y = layer(x, ei)
        sd = layer.state_dict()
        cases[name + "_x"] = x.numpy()
        cases[name + "_y"] = y.numpy()
        cases[name + "_W"] = np.stack([sd[f"heads.{h}.W.weight"].numpy() for h in range(heads)])
        cases[name + "_a"] = np.stack([sd[f"heads.{h}.a.weight"].numpy().reshape(-1) for h in range(heads)])
    cases["ei"] = ei.numpy()
    np.savez(os.path.join(HERE, "layers_random_graph.npz"), **cases)

    # ---- N-cut: reference MinCutRefinement with a GAT predictor ---------------------
    torch.manual_seed(5)
    hp, wp = 6, 5
    hfeat = 0.25 * torch.randn(hp * wp, 64, generator=g)          # small scale so weights do not underflow
    _, ei = pgc.construct_patch_graph(torch.zeros(1, hp * 16, wp * 16), hfeat)
    pred = R.GATNetwork(64, 32, 3, 2, 1, 0.1, 0.2).eval()
    mc = R.MinCutRefinement().eval()
    with torch.no_grad():
        loss, S = mc(hfeat, ei, 3, pred)
        w = mc.compute_edge_weights_for_ncut(hfeat, ei)
        # unit-scale features: every weight underflows -> the reference returns python 0.0
        big = 6.0 * torch.randn(hp * wp, 64, generator=g)
        loss0, _ = mc(big, ei, 3, pred)
    W, a = sd_stacks(pred)
    np.savez(os.path.join(HERE, "ncut.npz"), h=hfeat.numpy(), grid=np.array([hp, wp]), W=W, a=a,
             S=S.numpy(), loss=np.float32(float(loss)), w=w.numpy(),
             h_big=big.numpy(), loss_big=np.float32(float(loss0)))

    # ---- patch extraction + mean pooling ---------------------------------------------
    fm = torch.randn(5, 70, 75, generator=g)
    patches, (nph, npw) = pgc.image_to_patches(fm)
    np.savez(os.path.join(HERE, "patch_pool.npz"), fm=fm.numpy(), grid=np.array([nph, npw]),
             pooled=patches.mean(dim=(2, 3)).numpy(), patches_sha=np.array(sha16(patches.numpy())),
             scalar_mean=patches.mean(dim=[1, 2, 3]).numpy())

    # ---- whole block, one image at a time, as train_end_to_end.py:318-421 ---------------
    blk = {}
    for tag, (H, Wd, in_dim, K, scale) in {
        "64x64": (64, 64, 20, 2, 1.0),
        "128x96": (128, 96, 20, 2, 1.0),
        "70x75": (70, 75, 12, 3, 1.0),          # padded grid, non-divisible nearest un-pool, K=3
        "256x256": (256, 256, 20, 2, 1.0),      # BASELINE config 1
    }.items():
        torch.manual_seed(1234)
        patch_gat = R.GATNetwork(in_dim, 128, 64, 4, 1, 0.1, 0.2).eval()
        if R.PatchSegmentPredictor is not None:
            predictor = R.PatchSegmentPredictor(64, K, 32, use_gnn=True, num_gnn_layers=1, num_heads=2).eval()
            pred_net = predictor.gnn_predictor
        else:
            predictor = pred_net = R.GATNetwork(64, 32, K, 2, 1, 0.1, 0.2).eval()
        region_gat = R.GATNetwork(64, 128, 64, 4, 1, 0.1, 0.2).eval()
        mcm = R.MinCutRefinement().eval()
        img = torch.zeros(3, H, Wd)
        with torch.no_grad():
            _, (nph, npw) = pgc.image_to_patches(img)                          # :318
            x = scale * torch.randn(nph * npw, in_dim, generator=g)            # (:326 placeholder)
            _, ei = pgc.construct_patch_graph(img, x)                          # :329
            h = patch_gat(x, ei)                                               # :332
            loss, S = mcm(h, ei, K, predictor)                                 # :348
            hard = torch.argmax(S, dim=1)                                      # :356
            Rf = torch.zeros(K, 64)                                            # :367-373
            for k in range(K):
                m = hard == k
                if m.sum() > 0:
                    Rf[k] = h[m].mean(dim=0)
            s_, t_ = torch.triu_indices(K, K, offset=1)                        # :376-378
            rei = torch.stack([torch.cat([s_, t_]), torch.cat([t_, s_])], 0)
            G = region_gat(Rf, rei)                                            # :384
            P = G[hard]                                                        # :404
            fg = F.interpolate(P.T.reshape(64, nph, npw).unsqueeze(0), size=(H, Wd), mode="nearest").squeeze(0)  # :411-421
        for nm, net in (("patch", patch_gat), ("pred", pred_net), ("region", region_gat)):
            W, a = sd_stacks(net)
            blk[f"{tag}_{nm}_W"], blk[f"{tag}_{nm}_a"] = W, a
        blk[f"{tag}_x"] = x.numpy()
        blk[f"{tag}_meta"] = np.array([H, Wd, in_dim, K, nph, npw])
        blk[f"{tag}_h"] = h.numpy()
        blk[f"{tag}_S"] = S.numpy()
        blk[f"{tag}_loss"] = np.float32(float(loss))
        blk[f"{tag}_hard"] = hard.numpy()
        blk[f"{tag}_R"] = Rf.numpy()
        blk[f"{tag}_G"] = G.numpy()
        blk[f"{tag}_rei"] = rei.numpy()
        blk[f"{tag}_fg_sha"] = np.array(sha16(fg.numpy()))
        blk[f"{tag}_fg_sum"] = np.float64(fg.double().sum().item())
        if H * Wd <= 70 * 75:
            blk[f"{tag}_fg"] = fg.numpy()
    np.savez_compressed(os.path.join(HERE, "block_images.npz"), **blk)

    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
