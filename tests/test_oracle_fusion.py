"""CPU: FeatureFusion (model/fusion_detection/feature_fusion.py, the consumer of the block's output) — the oracle
restatement against the fixtures of the untouched reference (tests/golden/make_golden_fusion.py) and the live reference
when mounted; the error behaviour of the B200 module; and a thread-by-thread Python model of
``region_map_gather_vec_kernel``'s index arithmetic (csrc/fusion.cu), which was written without access to a GPU."""
import importlib
import warnings

import numpy as np
import pytest
import torch

from oracle import ref_loader, restate as O

REGION_CASES = ["rand_i64", "rand_i32", "invalid", "blocky", "odd", "allbad"]


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("tag", REGION_CASES)
def test_region_branch_matches_reference_fixture(golden, tag):
    g = golden("fusion.npz")
    fu, table, m, ref = T(g[f"rg_{tag}_fu"]), T(g[f"rg_{tag}_table"]), T(g[f"rg_{tag}_map"]), T(g[f"rg_{tag}_out"])
    out = O.feature_fusion([fu], table, region_to_pixel_map=m)
    assert out.dtype == torch.float32 and torch.equal(out, ref)                # a gather: bit for bit
    assert torch.equal(O.region_map_gather(table, m), ref[:, fu.shape[1]:])
    if tag == "allbad":
        assert float(ref[:, fu.shape[1]:].abs().max()) == 0.0
    if tag == "invalid":
        bad = (m < 0) | (m >= table.shape[0])
        assert bad.any() and float(ref[:, fu.shape[1]:].permute(0, 2, 3, 1)[bad].abs().max()) == 0.0


def test_dense_branches_and_add_match_reference_fixture(golden):
    g = golden("fusion.npz")
    out = O.feature_fusion([T(g["d4_same_fu"])], T(g["d4_same_fg"]), target_spatial_size=(16, 16))
    assert torch.equal(out, T(g["d4_same_out"]))
    out = O.feature_fusion([T(g["d4_resize_fu0"]), T(g["d4_resize_fu1"])], T(g["d4_resize_fg"]))
    assert out.shape == (2, 4 + 3 + 8, 16, 24) and torch.allclose(out, T(g["d4_resize_out"]), atol=1e-6, rtol=0)
    out = O.feature_fusion([T(g["add_fu"])], T(g["add_table"]), region_to_pixel_map=T(g["add_map"]), fusion_method="add")
    assert torch.equal(out, T(g["add_out"]))


def _ref_ff():
    ref_loader.load()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module("model.fusion_detection.feature_fusion").FeatureFusion


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_oracle_and_module_errors_match_live_reference():
    import mingraph_unet_b200 as mg
    Ref = _ref_ff()
    gen = torch.Generator().manual_seed(5)
    fu, table = torch.randn(2, 3, 8, 12, generator=gen), torch.randn(6, 10, generator=gen)
    m = torch.randint(-2, 9, (2, 8, 12), generator=gen)
    assert torch.equal(Ref([3], 10)([fu], table, region_to_pixel_map=m), O.feature_fusion([fu], table, region_to_pixel_map=m))
    # same exceptions, raised before any device work (so they are testable without a GPU)
    cases = [
        (dict(fusion_method="concat"), dict(f_g=table), ValueError),                          # 2-D f_g without a map (:139-141)
        (dict(fusion_method="concat"), dict(f_g=torch.zeros(2, 3, 4)), ValueError),           # 3-D f_g
        (dict(fusion_method="add"), dict(f_g=torch.zeros(2, 5, 8, 12)), ValueError),          # channel mismatch (:146-147)
        (dict(fusion_method="multiply"), dict(f_g=torch.zeros(2, 3, 8, 12)), NotImplementedError),   # (:150)
    ]
    for ctor, call, exc in cases:
        for cls in (Ref, mg.FeatureFusion):
            with pytest.raises(exc) as ei:
                cls([3], 10, **ctor)([fu], **call)
            if cls is Ref:
                msg = str(ei.value)
            else:
                assert str(ei.value) == msg
        with pytest.raises(exc):
            O.feature_fusion([fu], fusion_method=ctor["fusion_method"], **call)
    # no CPU fallback in the product
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        mg.FeatureFusion([3], 10)([fu], table, region_to_pixel_map=m)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        mg.ops.region_map_gather(table, m)


# ---------------------------------------------------------------------------------------------
# thread-by-thread model of region_map_gather_vec_kernel (csrc/fusion.cu): same grid / block decomposition, same
# offset expressions, writing into a flat output buffer with a batch stride (channel slice of a fusion buffer)
# ---------------------------------------------------------------------------------------------
kFuTX, kFuTY, kFuRY = 64, 4, 16


def _model_vec_kernel(table, rmap, VEC, c_total, c0, dchunk):
    R, D = table.shape
    B, H, W = rmap.shape
    assert W % VEC == 0 and D % 4 == 0 and dchunk % 4 == 0
    buf = np.full(B * c_total * H * W, np.nan, dtype=np.float32)
    writes = np.zeros_like(buf, dtype=np.int32)
    out_off, stride, plane = c0 * H * W, c_total * H * W, H * W           # out pointer = buf + out_off
    tab, mp = table.reshape(-1), rmap.reshape(-1)
    dchunks = -(-D // dchunk)
    gx, gy, gz = -(-(W // VEC) // kFuTX), -(-H // kFuRY), B * dchunks
    for bz in range(gz):
        b, d0 = bz // dchunks, (bz - (bz // dchunks) * dchunks) * dchunk
        nd4 = min(dchunk, D - d0) >> 2
        for by in range(gy):
            for bx in range(gx):
                for ty in range(kFuTY):
                    for tx in range(kFuTX):
                        xv = bx * kFuTX + tx
                        if xv * VEC >= W:
                            continue
                        yend = min(H, (by + 1) * kFuRY)
                        for y in range(by * kFuRY + ty, yend, kFuTY):
                            moff = b * H * W + y * W + xv * VEC
                            lab = [int(l) if 0 <= l < R else -1 for l in mp[moff:moff + VEC]]
                            orow = out_off + b * stride + (d0 * H + y) * W + xv * VEC
                            for d4 in range(nd4):
                                for dd in range(4):
                                    vals = [tab[l * D + d0 + 4 * d4 + dd] if l >= 0 else 0.0 for l in lab]
                                    o = orow + (4 * d4 + dd) * plane
                                    buf[o:o + VEC] = vals
                                    writes[o:o + VEC] += 1
    return buf.reshape(B, c_total, H, W), writes.reshape(B, c_total, H, W)


@pytest.mark.parametrize("VEC,B,H,W,R,D,dchunk,c_total,c0", [
    (4, 2, 18, 24, 5, 8, 8, 11, 3),          # fp32 pack, ragged last row block (18 = 16 + 2), one channel chunk
    (8, 1, 5, 16, 3, 64, 32, 70, 6),         # bf16 pack, two channel chunks of 32
    (4, 2, 33, 264, 6, 36, 32, 36, 0),       # W/VEC = 66 > kFuTX: two x blocks (the second mostly idle); chunks 32 + 4
])
def test_vec_kernel_index_model_covers_the_slice_exactly_once(VEC, B, H, W, R, D, dchunk, c_total, c0):
    rng = np.random.default_rng(VEC * 1000 + W)
    table = rng.standard_normal((R, D)).astype(np.float32)
    rmap = rng.integers(-1, R + 2, size=(B, H, W))
    buf, writes = _model_vec_kernel(table, rmap, VEC, c_total, c0, dchunk)
    ref = O.region_map_gather(T(table), T(rmap)).numpy()
    assert np.array_equal(buf[:, c0:c0 + D], ref)                           # every element of the slice, right value
    assert np.all(writes[:, c0:c0 + D] == 1)                                # written exactly once
    assert np.all(writes[:, :c0] == 0) and np.all(writes[:, c0 + D:] == 0)  # nothing outside the slice
