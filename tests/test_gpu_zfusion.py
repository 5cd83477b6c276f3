"""GPU parity of FeatureFusion / mg_region_map_gather (csrc/fusion.cu) through the C ABI against the fixtures of the
untouched reference (tests/golden/fusion.npz) and the CPU oracle: the per-region gather is bit-exact in fp32 (it moves
values), bf16 output equals the rounded fp32 result, the dense branches equal torch's own cat / bilinear resize.

First run on hardware in round 2 (12 passed; rates in profiles/r2_fusion_gather.md).  On the CPU the kernel SOURCE is
also compiled for the host and executed thread by thread (tests/test_fusion_emulation.py), and
tests/test_oracle_fusion.py models its index arithmetic."""

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu]

from oracle import restate as O  # noqa: E402

REGION_CASES = ["rand_i64", "rand_i32", "invalid", "blocky", "odd", "allbad"]


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope="module")
def M():
    import mingraph_unet_b200 as m
    return m


@pytest.mark.parametrize("tag", REGION_CASES)
def test_region_branch_golden(golden, M, tag):
    g = golden("fusion.npz")
    fu, table, m, ref = T(g[f"rg_{tag}_fu"]), T(g[f"rg_{tag}_table"]), T(g[f"rg_{tag}_map"]), T(g[f"rg_{tag}_out"])
    mod = M.FeatureFusion([fu.shape[1]], table.shape[1])
    out = mod([fu.cuda()], table.cuda(), region_to_pixel_map=m.cuda())
    assert out.dtype == torch.float32 and torch.equal(out.cpu(), ref)
    # bf16 storage of the fused map (an addition): the rounded fp32 result
    buf = torch.empty(ref.shape, dtype=torch.bfloat16, device="cuda")
    out16 = mod([fu.cuda().bfloat16()], table.cuda(), region_to_pixel_map=m.cuda(), out=buf)
    assert out16.data_ptr() == buf.data_ptr()
    assert torch.equal(out16[:, fu.shape[1]:].cpu(), ref[:, fu.shape[1]:].bfloat16())
    assert torch.equal(out16[:, :fu.shape[1]].cpu(), fu.bfloat16())


def test_dense_branches_and_add_golden(golden, M):
    g = golden("fusion.npz")
    out = M.FeatureFusion([6], 10)([T(g["d4_same_fu"]).cuda()], T(g["d4_same_fg"]).cuda(), target_spatial_size=(16, 16))
    assert torch.equal(out.cpu(), T(g["d4_same_out"]))
    out = M.FeatureFusion([4, 3], 8)([T(g["d4_resize_fu0"]).cuda(), T(g["d4_resize_fu1"]).cuda()], T(g["d4_resize_fg"]).cuda())
    assert float((out.cpu() - T(g["d4_resize_out"])).abs().max()) <= 1e-5          # CUDA vs CPU bilinear resize
    out = M.FeatureFusion([16], 16, "add")([T(g["add_fu"]).cuda()], T(g["add_table"]).cuda(),
                                           region_to_pixel_map=T(g["add_map"]).cuda())
    assert torch.equal(out.cpu(), T(g["add_out"]))


def test_inputs_already_in_place_are_not_copied(M):
    """Scope row f1: the block's un-pool writes F_g into fused[:, C_u:], the decoder F_u into fused[:, :C_u]; the fusion
    is then free."""
    B, Cu, D, H, W = 2, 32, 64, 64, 48
    fused = torch.randn(B, Cu + D, H, W, device="cuda")
    keep = fused.clone()
    out = M.FeatureFusion([Cu], D)([fused[:, :Cu]], fused[:, Cu:], out=fused)
    assert out.data_ptr() == fused.data_ptr() and torch.equal(out, keep)


@pytest.mark.parametrize("odt,mdt", [(torch.float32, torch.int64), (torch.bfloat16, torch.int32), (torch.bfloat16, torch.int64)])
def test_region_gather_full_size_equals_torch_embedding(M, odt, mdt):
    """cfg-2 size (16 x 512 x 512, D = 64) straight into the channel slice [32:96] of a fusion buffer: equals the torch
    embedding gather bit for bit, and the neighbouring channels are untouched."""
    B, H, W, R, D, Cu = 16, 512, 512, 300, 64, 32
    gen = torch.Generator().manual_seed(9)
    table = torch.randn(R, D, generator=gen).cuda()
    cells = torch.randint(-1, R + 1, (B, H // 8, W // 4), generator=gen)           # runs of equal labels and invalid ones
    m = cells.repeat_interleave(8, 1).repeat_interleave(4, 2).to(mdt).cuda()
    fused = torch.full((B, Cu + D + 4, H, W), 7.0, dtype=odt, device="cuda")
    M.ops.region_map_gather(table, m, out=fused[:, Cu:Cu + D])
    idx = m.long()
    ok = (idx >= 0) & (idx < R)
    ref = torch.where(ok.unsqueeze(-1), table[idx.clamp(0, R - 1)], torch.zeros((), device="cuda")).permute(0, 3, 1, 2)
    assert torch.equal(fused[:, Cu:Cu + D], ref.to(odt))
    assert bool((fused[:, :Cu] == 7.0).all()) and bool((fused[:, Cu + D:] == 7.0).all())


def test_region_gather_unaligned_shapes_take_the_scalar_kernel(M):
    gen = torch.Generator().manual_seed(10)
    table = torch.randn(11, 6, generator=gen).cuda()
    m = torch.randint(-1, 12, (3, 9, 13), generator=gen).cuda()
    for odt in (torch.float32, torch.bfloat16):
        out = M.ops.region_map_gather(table, m, out_dtype=odt)
        assert torch.equal(out.cpu(), O.region_map_gather(table.cpu(), m.cpu()).to(odt))
    with pytest.raises(RuntimeError):
        M.ops.region_map_gather(table, m.float())
    with pytest.raises(ValueError):
        M.ops.region_map_gather(table, m, out=torch.empty(3, 6, 9, 14, device="cuda"))
