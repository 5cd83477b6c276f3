"""CPU: size-independent properties of the graph block, checked on the oracle with hypothesis-drawn shapes (the GPU
parity tests check the CUDA path against this oracle; these pin the oracle's own algebra beyond the fixed fixtures)."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import cint, restate as O

FAST = settings(max_examples=25, deadline=None)


@FAST
@given(st.integers(1, 12), st.integers(1, 12))
def test_grid_graph_structure(hp, wp):
    """patch_graph_construction.py:78-97: E = 2(Hp(Wp-1)+Wp(Hp-1)), symmetric, no self loops, no duplicates,
    in-degree 2/3/4 by position; numpy and C restatements agree."""
    e = O.grid_edge_index(hp, wp)
    assert e.shape == (2, 2 * (hp * (wp - 1) + wp * (hp - 1))) and np.array_equal(e, cint.grid_edge_index(hp, wp))
    if e.shape[1] == 0:
        return
    pairs = set(map(tuple, e.T))
    assert len(pairs) == e.shape[1] and all((t, s) in pairs for s, t in pairs) and all(s != t for s, t in pairs)
    deg = np.bincount(e[1], minlength=hp * wp).reshape(hp, wp)
    want = np.full((hp, wp), 4)
    want[0, :] -= 1; want[-1, :] -= 1; want[:, 0] -= 1; want[:, -1] -= 1
    assert np.array_equal(deg, want)
    # every edge joins 4-neighbours
    r0, c0 = np.divmod(e[0], wp); r1, c1 = np.divmod(e[1], wp)
    assert np.all(np.abs(r0 - r1) + np.abs(c0 - c1) == 1)


@FAST
@given(st.integers(2, 40), st.integers(1, 6), st.integers(1, 3), st.integers(0, 10 ** 6))
def test_gat_layer_is_permutation_equivariant_and_edge_order_invariant(n, fin, heads, seed):
    """graph_attention.py:53-118: relabelling the nodes permutes the output rows; the order of the edge list does not
    matter (softmax shift is a global max, sums are over incoming edges) — up to fp32 summation order."""
    g = torch.Generator().manual_seed(seed)
    fout = 4
    x = torch.randn(n, fin, generator=g)
    E = 3 * n
    ei = torch.randint(0, n, (2, E), generator=g)
    Ws, As = torch.randn(heads, fout, fin, generator=g), torch.randn(heads, 2 * fout, generator=g)
    y = O.gat_layer(x, ei, Ws, As, 0.2, concat=False)
    perm = torch.randperm(n, generator=g)
    inv = torch.empty_like(perm); inv[perm] = torch.arange(n)
    y_perm = O.gat_layer(x[perm], inv[ei], Ws, As, 0.2, concat=False)          # node i of the new graph = node perm[i]
    assert torch.allclose(y_perm, y[perm], atol=2e-5)
    shuffle = torch.randperm(E, generator=g)
    assert torch.allclose(O.gat_layer(x, ei[:, shuffle], Ws, As, 0.2, concat=False), y, atol=2e-5)
    # rows without incoming edges are exactly zero (ELU(0) = 0)
    indeg = torch.bincount(ei[1], minlength=n)
    assert torch.all(y[indeg == 0] == 0)
    # concat = the same heads side by side
    yc = O.gat_layer(x, ei, Ws, As, 0.2, concat=True)
    assert torch.allclose(yc.view(n, heads, fout).mean(1), y, atol=1e-6)


@FAST
@given(st.integers(2, 8), st.integers(2, 8), st.integers(2, 4), st.integers(0, 10 ** 6))
def test_ncut_loss_properties(hp, wp, K, seed):
    """mincut_refinement.py:92-152: a one-segment partition cuts nothing; the loss is non-negative, at most K, and does
    not depend on the edge order."""
    g = torch.Generator().manual_seed(seed)
    n = hp * wp
    h = 0.3 * torch.randn(n, 8, generator=g)
    ei = torch.from_numpy(O.grid_edge_index(hp, wp))
    one = torch.zeros(n, K); one[:, 0] = 1.0
    assert float(O.ncut_loss(h, ei, one, K)) == 0.0
    S = torch.softmax(torch.randn(n, K, generator=g), 1)
    loss = float(O.ncut_loss(h, ei, S, K))
    assert 0.0 <= loss <= K + 1e-5
    sh = torch.randperm(ei.shape[1], generator=g)
    assert abs(float(O.ncut_loss(h, ei[:, sh], S, K)) - loss) <= 1e-5 * max(1.0, loss)
    # weights are in (0, 1] and symmetric on the symmetric grid graph
    w = O.ncut_edge_weights(h, ei)
    assert torch.all(w > 0) and torch.all(w <= 1)
    wd = {(int(s), int(t)): float(v) for s, t, v in zip(ei[0], ei[1], w)}
    assert all(wd[(t, s)] == v for (s, t), v in wd.items())


@FAST
@given(st.integers(1, 6), st.integers(1, 6), st.integers(1, 5), st.integers(0, 10 ** 6))
def test_unpool_then_pool_round_trip(nph, npw, D, seed):
    """train_end_to_end.py:404-421 with divisible sizes: every patch value is replicated 16x16, so the patch mean of the
    un-pooled map returns the per-patch rows (un-pool -> pool = identity), numpy/torch and C un-pool agree bit for bit."""
    g = torch.Generator().manual_seed(seed)
    P = torch.randn(nph * npw, D, generator=g)
    H, W = nph * 16, npw * 16
    dense = O.unpool_nearest(P, nph, npw, H, W)
    assert np.array_equal(dense.numpy(), cint.unpool_nearest(P.numpy(), None, nph, npw, H, W))
    assert torch.allclose(O.patch_mean_pool(dense, 16), P, atol=1e-6)
    assert torch.equal(dense[:, ::16, ::16].reshape(D, -1).t(), P)


@FAST
@given(st.integers(6, 40), st.integers(1, 5), st.integers(1, 6), st.integers(0, 10 ** 6))
def test_knn_oracle_properties(n, k, d, seed):
    """kNN restatement (not in the reference): k distinct non-self neighbours per node, sorted by (distance, id), and
    no excluded node is closer than the k-th neighbour."""
    rng = np.random.default_rng(seed)
    x = rng.integers(-2, 3, size=(n, d)).astype(np.float32)           # small integers: many exact ties
    ei, dist = O.knn_graph(x, k)
    src = ei[0].reshape(n, k)
    full = O.knn_sqdist(x)
    for i in range(n):
        assert i not in src[i] and len(set(src[i])) == k
        keys = [(float(full[i, j]), int(j)) for j in src[i]]
        assert keys == sorted(keys) and np.array_equal(dist[i], full[i, src[i]])
        rest = [(float(full[i, j]), j) for j in range(n) if j != i and j not in src[i]]
        assert all(r > keys[-1] for r in rest)
    assert np.array_equal(ei[1], np.repeat(np.arange(n), k))


@FAST
@given(st.integers(1, 40), st.integers(1, 6), st.integers(0, 10 ** 6))
def test_region_mean_pool_properties(n, K, seed):
    """train_end_to_end.py:368-373: empty regions give zero rows; count-weighted region means recover the global mean."""
    g = torch.Generator().manual_seed(seed)
    h = torch.randn(n, 5, generator=g)
    hard = torch.randint(0, K, (n,), generator=g)
    R = O.region_mean_pool(h, hard, K)
    cnt = torch.bincount(hard, minlength=K).float()
    assert torch.all(R[cnt == 0] == 0)
    assert torch.allclose((R * cnt[:, None]).sum(0) / n, h.mean(0), atol=1e-5)
