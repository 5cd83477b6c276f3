"""CPU: the oracle restatement against the fixtures produced by the untouched
reference (tests/golden/make_golden.py) and, when mounted, the live reference."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import ref_loader, restate as O


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def T(a):
    return torch.from_numpy(np.asarray(a))


def same(a, b, atol=2e-6):
    """The restatement repeats the reference's op order, so it is bitwise equal
    under the same BLAS threading; across thread counts only matmul blocking
    differs (observed <= 2e-7), hence a tight absolute bound instead of equality."""
    return a.shape == b.shape and float((a - b).abs().max() if a.numel() else 0.0) <= atol


def test_kat1_analytic_head(golden):
    g = golden("kat1_head.npz")
    y = O.gat_head(T(g["x"]), T(g["ei"]), torch.eye(2), torch.tensor([1.0, 0, 0, 1.0]))
    assert torch.equal(y, T(g["y"]))
    # hand-derived: e=(3,0,-0.2,0), M=3
    np.testing.assert_allclose(y.numpy(), [[-0.3624776602, 0.6495020390], [1, 0], [1, 0]], atol=1e-6)


def test_kat2_default_patch_gat(golden):
    g = golden("kat2_patch_gat.npz")
    ei = T(O.grid_edge_index(16, 16))
    assert sha16(ei.numpy()) == str(g["ei_sha"]) == "79180fa641eb7814"
    y = O.gat_network(T(g["x"]), ei, T(g["W"]), T(g["a"]))
    assert same(y, T(g["y"]))
    assert abs(float(y.sum()) - 1189.990723) < 2e-2
    np.testing.assert_allclose(y[0, :3].numpy(), [-0.1591568, 0.4012290, -0.2805685], atol=1e-5)


def test_kat3_edge_index_all_shapes(golden):
    g = golden("kat3_edge_index.npz")
    shapes = [k[4:] for k in g.files if k.startswith("sha_")]
    assert len(shapes) >= 12
    for s in shapes:
        hp, wp = map(int, s.split("x"))
        e = O.grid_edge_index(hp, wp)
        assert e.dtype == np.int64 and tuple(e.shape) == tuple(g[f"shape_{s}"])
        assert sha16(e) == str(g[f"sha_{s}"])
        if f"ei_{s}" in g.files:
            assert np.array_equal(e, g[f"ei_{s}"])
            assert np.array_equal(O.grid_edge_index_loop(hp, wp), g[f"ei_{s}"])
    assert str(g["sha_32x32"]) == "117dbf3e9444f2fc" and str(g["sha_64x64"]) == "8ba2cf2fc8a5b476"
    assert O.grid_edge_index(1, 1).shape == (2, 0)


@pytest.mark.parametrize("name,concat", [("avg_17_24_3", False), ("cat_33_32_4", True),
                                         ("avg_64_2_2", False), ("avg_130_40_1", False)])
def test_layers_random_multigraph(golden, name, concat):
    g = golden("layers_random_graph.npz")
    y = O.gat_layer(T(g[name + "_x"]), T(g["ei"]), T(g[name + "_W"]), T(g[name + "_a"]), 0.2, concat)
    assert same(y, T(g[name + "_y"]))
    assert torch.all(y[5] == 0) and torch.all(y[-3:] == 0)      # zero in-degree rows are exactly 0


def test_ncut(golden):
    g = golden("ncut.npz")
    hp, wp = g["grid"]
    ei = T(O.grid_edge_index(int(hp), int(wp)))
    h = T(g["h"])
    assert torch.equal(O.ncut_edge_weights(h, ei), T(g["w"]))
    pred = lambda f, e: O.gat_network(f, e, T(g["W"]), T(g["a"]))
    loss, S = O.mincut_forward(h, ei, 3, pred)
    assert same(S, T(g["S"]))
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-5) and float(loss) > 0
    loss0, _ = O.mincut_forward(T(g["h_big"]), ei, 3, pred)
    assert float(loss0) == float(g["loss_big"]) == 0.0          # all weights underflow -> python 0.0 in the reference
    with pytest.raises(ValueError):
        O.ncut_loss(h, ei, S[:, :2], 3)


def test_patch_pool(golden):
    g = golden("patch_pool.npz")
    fm = T(g["fm"])
    p, grid = O.image_to_patches(fm)
    assert grid == tuple(g["grid"]) == (5, 5) and sha16(p.numpy()) == str(g["patches_sha"])
    assert torch.equal(O.patch_mean_pool(fm), T(g["pooled"]))


@pytest.mark.parametrize("tag", ["64x64", "128x96", "70x75", "256x256"])
def test_block_per_image(golden, tag):
    g = golden("block_images.npz")
    H, W, in_dim, K, nph, npw = (int(v) for v in g[f"{tag}_meta"])
    params = {f"{n}_{p}": T(g[f"{tag}_{n}_{p}"]) for n in ("patch", "pred", "region") for p in ("W", "a")}
    r = O.graph_block_image(T(g[f"{tag}_x"]), H, W, params, K=K)
    assert r["grid"] == (nph, npw)
    assert same(r["h"], T(g[f"{tag}_h"]))
    assert same(r["S"], T(g[f"{tag}_S"]))
    assert torch.equal(r["hard"], T(g[f"{tag}_hard"]))
    assert float(r["loss"]) == pytest.approx(float(g[f"{tag}_loss"]), rel=1e-5)
    assert same(r["region_in"], T(g[f"{tag}_R"]))
    assert same(r["region_out"], T(g[f"{tag}_G"]))
    assert np.array_equal(O.complete_edge_index(K), g[f"{tag}_rei"])
    assert abs(float(r["f_g"].double().sum()) - float(g[f"{tag}_fg_sum"])) <= 1e-6 * r["f_g"].numel()
    if f"{tag}_fg" in g.files:
        assert same(r["f_g"], T(g[f"{tag}_fg"]))


@pytest.mark.parametrize("o,i", [(64, 4), (70, 5), (75, 5), (130, 9), (33, 33), (66, 33), (1000, 63), (17, 2)])
def test_nearest_index_matches_torch(o, i):
    src = torch.arange(i, dtype=torch.float32).view(1, 1, i, 1)
    got = torch.nn.functional.interpolate(src, size=(o, 1), mode="nearest").view(-1).long().numpy()
    assert np.array_equal(O.nearest_index(o, i), got)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_live_reference_agrees():
    R = ref_loader.load()
    torch.manual_seed(3)
    net = R.GATNetwork(20, 128, 64, 4, 1, 0.1, 0.2).eval()
    x = torch.randn(15 * 9, 20)
    _, ei = R.PatchGraphConstructor(16).construct_patch_graph(torch.zeros(3, 15 * 16 - 1, 9 * 16), x)
    assert np.array_equal(ei.numpy(), O.grid_edge_index(15, 9))
    Ws, As = O.stack_from_state_dict(net.state_dict())
    with torch.no_grad():
        assert same(net(x, ei), O.gat_network(x, ei, Ws, As))


def test_knn_oracle_matches_pure_python_definition():
    """The vectorised kNN restatement equals a literal per-pair loop (sequential fp32 sub/mul/add, lexicographic
    (distance, id) order, self excluded) — including exact ties from duplicated points."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((24, 5)).astype(np.float32)
    x[7] = x[3]                      # duplicate point -> zero distance, ties
    x[11] = x[3]
    k = 4
    ei, dist = O.knn_graph(x, k, nodes_per_graph=12)
    assert ei.shape == (2, 24 * k) and ei.dtype == np.int64
    for j in range(24):
        g0 = (j // 12) * 12
        cands = []
        for i in range(g0, g0 + 12):
            if i == j:
                continue
            acc = np.float32(0)
            for d in range(5):
                t = np.float32(x[j, d] - x[i, d])
                acc = np.float32(acc + np.float32(t * t))
            cands.append((acc, i))
        cands.sort()
        want = cands[:k]
        assert [int(v) for v in ei[0, j * k:(j + 1) * k]] == [i for _, i in want]
        assert all(int(t) == j for t in ei[1, j * k:(j + 1) * k])
        assert [float(v) for v in dist[j]] == [float(d) for d, _ in want]


# ---- scope row f4: losses and the multi-layer stack ------------------------------------------------
FL_CASES = ["b3_n37_d64", "b1_n256_d64", "b2_n50_d7"]
TV_CASES = ["b2_c1_64x64", "b3_c2_37x53", "b1_c3_5x200", "b2_c2_19x8"]


@pytest.mark.parametrize("tag", FL_CASES)
def test_feature_consistency_loss_restatement(golden, tag):
    g = golden("losses.npz")
    fu = T(g[f"fl_{tag}_fu"]).requires_grad_(True)
    fg = T(g[f"fl_{tag}_fg"]).requires_grad_(True)
    loss = O.feature_consistency_loss(fu, fg, T(g[f"fl_{tag}_y"]), float(g[f"fl_{tag}_margin"]))
    loss.backward()
    assert float(loss) == pytest.approx(float(g[f"fl_{tag}_loss"]), rel=1e-6)
    assert same(fu.grad, T(g[f"fl_{tag}_gfu"]), 1e-6) and same(fg.grad, T(g[f"fl_{tag}_gfg"]), 1e-6)


@pytest.mark.parametrize("tag", TV_CASES)
def test_tv_loss_restatement(golden, tag):
    g = golden("losses.npz")
    x = T(g[f"tv_{tag}_x"]).requires_grad_(True)
    loss = O.tv_loss(x, float(g[f"tv_{tag}_weight"]))
    loss.backward()
    assert float(loss) == pytest.approx(float(g[f"tv_{tag}_loss"]), rel=1e-6)
    assert same(x.grad, T(g[f"tv_{tag}_gx"]), 1e-7)


@pytest.mark.parametrize("tag", ["2layer", "3layer"])
def test_multilayer_stack_restatement(golden, tag):
    g = golden("multilayer_gat.npz")
    hp, wp, fin, hidden, fout, heads, nl = (int(v) for v in g[f"{tag}_meta"])
    layers = [(T(g[f"{tag}_W{i}"]), T(g[f"{tag}_a{i}"])) for i in range(nl)]
    y = O.gat_network_multilayer(T(g[f"{tag}_x"]), T(O.grid_edge_index(hp, wp)), layers)
    assert same(y, T(g[f"{tag}_y"]))
    assert "cannot be multiplied" in str(g[f"{tag}_ref_network_error"])    # the reference's own stack crashes


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_losses_against_live_reference():
    import importlib
    import warnings
    ref_loader.load()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fl = importlib.import_module("model.unet.feature_loss")
        te = importlib.import_module("scripts.train_end_to_end")
    gen = torch.Generator().manual_seed(9)
    fu, fg = 0.1 * torch.randn(2, 33, 16, generator=gen), 0.1 * torch.randn(2, 33, 16, generator=gen)
    y = torch.randint(0, 2, (2, 33), generator=gen)
    assert torch.equal(O.feature_consistency_loss(fu, fg, y, 0.7), fl.FeatureConsistencyLoss(0.7)(fu, fg, y))
    x = torch.rand(2, 3, 17, 9, generator=gen)
    assert torch.equal(O.tv_loss(x, 1.5), te.TVLoss(1.5)(x))
