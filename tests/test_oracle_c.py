"""CPU: the plain-C integer oracle (oracle/restate_int.c) against the numpy restatement, torch, and the
fixtures generated from the untouched reference."""
import hashlib

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import cint, restate as O


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_c_grid_edge_index_matches_reference_fixtures(golden):
    g = golden("kat3_edge_index.npz")
    shapes = [k[4:] for k in g.files if k.startswith("sha_")]
    assert len(shapes) >= 12
    for s in shapes:
        hp, wp = map(int, s.split("x"))
        e = cint.grid_edge_index(hp, wp)
        assert e.dtype == np.int64 and tuple(e.shape) == tuple(g[f"shape_{s}"]) and sha16(e) == str(g[f"sha_{s}"])
        assert np.array_equal(e, O.grid_edge_index(hp, wp))
    assert sha16(cint.grid_edge_index(32, 32)) == "117dbf3e9444f2fc"
    assert cint.grid_edge_index(1, 1).shape == (2, 0)


@pytest.mark.parametrize("K", [1, 2, 3, 5, 8])
def test_c_complete_edge_index(K, golden):
    e = cint.complete_edge_index(K)
    assert np.array_equal(e, O.complete_edge_index(K))
    if K > 1:
        s, t = torch.triu_indices(K, K, offset=1)                            # train_end_to_end.py:376-378
        assert np.array_equal(e, torch.stack([torch.cat([s, t]), torch.cat([t, s])], 0).numpy())
    if K == 2:
        assert np.array_equal(e, golden("block_images.npz")["64x64_rei"])


def test_c_csr_is_a_stable_sort():
    rng = np.random.default_rng(3)
    N, E = 97, 1500
    ei = rng.integers(0, N, size=(2, E)).astype(np.int64)
    ei[:, 10] = ei[:, 11]                                                    # duplicate edge
    for by_target in (True, False):
        key, val = (ei[1], ei[0]) if by_target else (ei[0], ei[1])
        order = np.argsort(key, kind="stable")
        rowptr, col, eid, bad = cint.csr_from_coo(ei, N, by_target)
        assert bad == 0 and np.array_equal(eid, order.astype(np.int32)) and np.array_equal(col, val[order].astype(np.int32))
        assert np.array_equal(rowptr, np.concatenate([[0], np.cumsum(np.bincount(key, minlength=N))]).astype(np.int32))
    ei[1, 5] = N + 3
    assert cint.csr_from_coo(ei, N, True)[3] == 1                            # out-of-range index is reported, not stored
    # the grid graph: CSR neighbour order = up, left, right, down for an interior node (ascending COO edge id)
    hp, wp = 5, 7
    rowptr, col, eid, _ = cint.csr_from_coo(cint.grid_edge_index(hp, wp), hp * wp, True)
    n = 2 * wp + 3
    assert list(col[rowptr[n]:rowptr[n + 1]]) == [n - wp, n - 1, n + 1, n + wp]


def test_c_argmax_first_maximum_like_torch():
    gen = torch.Generator().manual_seed(4)
    S = torch.softmax(torch.randn(500, 3, generator=gen), 1)
    S[7] = torch.tensor([0.25, 0.5, 0.5])                                    # tie: first maximum
    S[8] = torch.tensor([0.5, 0.5, 0.0])
    assert np.array_equal(cint.argmax_rows(S.numpy()), torch.argmax(S, 1).numpy().astype(np.int32))


@pytest.mark.parametrize("out_size,in_size", [(512, 32), (70, 5), (75, 5), (1000, 63), (33, 33), (64, 32), (7, 9), (1, 1)])
def test_c_nearest_index_matches_torch(out_size, in_size):
    ref = F.interpolate(torch.arange(in_size, dtype=torch.float32).view(1, 1, in_size), size=out_size, mode="nearest")
    idx = cint.nearest_index(out_size, in_size)
    assert np.array_equal(idx, ref.view(-1).numpy().astype(np.int32))
    assert np.array_equal(idx, O.nearest_index(out_size, in_size).astype(np.int32))


def test_c_unpool_matches_reference_fixture(golden):
    g = golden("block_images.npz")
    for tag in ("64x64", "70x75"):                                           # divisible and padded / non-divisible
        H, W, _, K, nph, npw = (int(v) for v in g[f"{tag}_meta"])
        out = cint.unpool_nearest(g[f"{tag}_G"], g[f"{tag}_hard"], nph, npw, H, W)
        # bit-exact against the reference's F.interpolate output and against torch run here
        assert sha16(out) == str(g[f"{tag}_fg_sha"]) and np.array_equal(out, g[f"{tag}_fg"])
        P = torch.from_numpy(g[f"{tag}_G"])[torch.from_numpy(g[f"{tag}_hard"])]
        assert np.array_equal(out, O.unpool_nearest(P, nph, npw, H, W).numpy())
