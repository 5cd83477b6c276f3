"""CPU: the device code of csrc/fusion_kernels.cuh compiled for the HOST with g++ (tests/emu/cuda_host_shim.h) and
executed block by block, thread by thread, with the launch shape csrc/fusion.cu chooses.  Unlike the Python model in
tests/test_oracle_fusion.py this runs the kernel SOURCE itself: the vector label loads, the float4 table reads, the bf16
packing, the ``same``-label fast path and every address expression — with loads and stores that check their alignment
(a misaligned 16-byte access faults on the GPU but not on x86) and stores that are bounds-checked against the output
allocation.  Results must equal the oracle bit for bit, the channels next to the target slice must stay untouched.

(The kernel was written without access to a GPU; this is the strongest check available without one.)"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle import restate as O

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
EMU = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.isfile(os.path.join(inc, "cuda_bf16.h")):
        pytest.skip("CUDA headers not available")
    so = str(tmp_path_factory.mktemp("emu") / "fusion_emu.so")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-w", "-I", inc, "-o", so,
                        os.path.join(EMU, "fusion_emu.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(so)
    lib.emu_region_map_gather.restype = C.c_int
    lib.emu_region_map_gather.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    return lib


def aligned(n_bytes, dtype, offset=0):
    """numpy array of ``dtype`` whose data starts ``offset`` bytes past a 64-byte boundary"""
    raw = np.zeros(n_bytes + 128, dtype=np.uint8)
    start = (-raw.ctypes.data) % 64 + offset
    return raw[start:start + n_bytes].view(dtype)


def run_case(lib, B, H, W, R, D, out_bf16, map_i64, c_total, c0, rng, sms=148, force_scalar=0, pattern="random",
             out_offset=0, expect_vec=None):
    table = aligned(R * D * 4, np.float32)
    table[:] = rng.standard_normal(R * D).astype(np.float32)
    mdt = np.int64 if map_i64 else np.int32
    m = aligned(B * H * W * np.dtype(mdt).itemsize, mdt)
    if pattern == "runs":
        cells = rng.integers(-1, R + 1, size=(B, -(-H // 4), -(-W // 8)))
        mm = np.repeat(np.repeat(cells, 4, axis=1), 8, axis=2)[:, :H, :W]
    else:
        mm = rng.integers(-2, R + 2, size=(B, H, W))
    m[:] = mm.reshape(-1)
    esz = 2 if out_bf16 else 4
    n_out = B * c_total * H * W
    buf = aligned(n_out * esz, np.uint16 if out_bf16 else np.float32, offset=out_offset)
    sentinel = 0x7FC1 if out_bf16 else np.float32(12345.0)
    buf[:] = sentinel
    base = buf.ctypes.data
    out_ptr = base + c0 * H * W * esz
    rc = lib.emu_region_map_gather(table.ctypes.data, R, D, m.ctypes.data, int(map_i64), B, H, W, out_ptr, int(out_bf16),
                                   c_total * H * W, base, base + n_out * esz, sms, force_scalar)
    assert rc >= 0, "the emulated kernel made a misaligned or out-of-range access"
    if expect_vec is not None:
        assert rc == int(expect_vec)
    ref = O.region_map_gather(torch.from_numpy(table.reshape(R, D).copy()), torch.from_numpy(mm.copy()))
    got = buf.reshape(B, c_total, H, W)
    if out_bf16:
        ref_bits = ref.bfloat16().view(torch.int16).numpy().view(np.uint16)
        assert np.array_equal(got[:, c0:c0 + D], ref_bits)
    else:
        assert np.array_equal(got[:, c0:c0 + D], ref.numpy())
    assert np.all(got[:, :c0] == sentinel) and np.all(got[:, c0 + D:] == sentinel)
    return rc


@pytest.mark.parametrize("out_bf16", [False, True])
@pytest.mark.parametrize("map_i64", [False, True])
def test_vector_kernel_runs_and_matches_the_oracle(emu, out_bf16, map_i64):
    rng = np.random.default_rng(10 * out_bf16 + map_i64)
    # ragged last row block (H = 21), more than one x block (W / VEC > 64 for fp32), channel slice in the middle
    run_case(emu, B=2, H=21, W=272, R=7, D=12, out_bf16=out_bf16, map_i64=map_i64, c_total=20, c0=5, rng=rng, expect_vec=True)
    # superpixel-like runs: the one-read-per-16-bytes path, invalid labels inside runs
    run_case(emu, B=1, H=16, W=64, R=5, D=64, out_bf16=out_bf16, map_i64=map_i64, c_total=64, c0=0, rng=rng, pattern="runs",
             expect_vec=True)


def test_channel_chunks_across_blockidx_z(emu):
    """few pixels on a big machine: channels are split into chunks of 32 across blockIdx.z (D = 72 -> 32 + 32 + 8)"""
    rng = np.random.default_rng(3)
    for out_bf16 in (False, True):
        run_case(emu, B=2, H=8, W=16, R=9, D=72, out_bf16=out_bf16, map_i64=True, c_total=80, c0=4, rng=rng, sms=148,
                 expect_vec=True)
        run_case(emu, B=2, H=8, W=16, R=9, D=72, out_bf16=out_bf16, map_i64=False, c_total=80, c0=4, rng=rng, sms=1,
                 expect_vec=True)       # a one-SM machine: all channels in one block


def test_unaligned_or_odd_shapes_take_the_scalar_kernel(emu):
    rng = np.random.default_rng(4)
    for out_bf16 in (False, True):
        run_case(emu, B=3, H=9, W=13, R=11, D=6, out_bf16=out_bf16, map_i64=True, c_total=9, c0=2, rng=rng, expect_vec=False)
        run_case(emu, B=1, H=4, W=16, R=3, D=6, out_bf16=out_bf16, map_i64=False, c_total=6, c0=0, rng=rng, expect_vec=False)
        # vector-friendly shape but the output pointer is only 4- / 2-byte aligned
        run_case(emu, B=1, H=4, W=16, R=3, D=8, out_bf16=out_bf16, map_i64=False, c_total=8, c0=0, rng=rng,
                 out_offset=2 if out_bf16 else 4, expect_vec=False)
        # and the scalar kernel on a vector-friendly shape
        run_case(emu, B=2, H=5, W=16, R=4, D=8, out_bf16=out_bf16, map_i64=True, c_total=10, c0=1, rng=rng, force_scalar=1,
                 expect_vec=False)


def test_golden_fixture_through_the_emulated_kernel(emu, golden):
    """the reference's own outputs (tests/golden/fusion.npz) reproduced by the emulated device code"""
    g = golden("fusion.npz")
    for tag in ["rand_i64", "rand_i32", "invalid", "blocky", "odd", "allbad"]:
        table, m, ref = g[f"rg_{tag}_table"], g[f"rg_{tag}_map"], g[f"rg_{tag}_out"]
        cu = g[f"rg_{tag}_fu"].shape[1]
        R, D = table.shape
        B, H, W = m.shape
        t = aligned(R * D * 4, np.float32); t[:] = table.reshape(-1)
        mm = aligned(m.size * m.dtype.itemsize, m.dtype); mm[:] = m.reshape(-1)
        buf = aligned(B * (cu + D) * H * W * 4, np.float32)
        buf.reshape(B, cu + D, H, W)[:, :cu] = g[f"rg_{tag}_fu"]
        base = buf.ctypes.data
        rc = emu.emu_region_map_gather(t.ctypes.data, R, D, mm.ctypes.data, int(m.dtype == np.int64), B, H, W,
                                       base + cu * H * W * 4, 0, (cu + D) * H * W, base, base + buf.nbytes, 148, 0)
        assert rc >= 0
        assert np.array_equal(buf.reshape(B, cu + D, H, W), ref), tag          # Concat(F_u, F_g) bit for bit


# ---------------------------------------------------------------------------------------------
# the Python side (modules.FeatureFusion -> ops.region_map_gather -> ctypes arguments) driven end to end on CPU tensors,
# with the C-ABI call routed to the emulated kernel: checks the wrapper's pointer / stride / dtype-code marshalling and
# the module's channel-slice bookkeeping against the reference fixtures
# ---------------------------------------------------------------------------------------------
class _NullCtx:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


@pytest.fixture
def routed(emu, monkeypatch):
    import mingraph_unet_b200 as mg
    calls = []

    def call(fn, *args):
        assert fn == "mg_region_map_gather"
        table, R, D, mp, map_dt, B, H, W, out, out_dt, stride, stream = args
        esz = 4 if out_dt == mg._lib.MG_F32 else 2
        hi = out + ((B - 1) * stride + D * H * W) * esz
        rc = emu.emu_region_map_gather(table, R, D, mp, int(map_dt == mg._lib.MG_I64), B, H, W, out,
                                       int(out_dt == mg._lib.MG_BF16), stride, out, hi, 148, 0)
        assert rc >= 0, "misaligned or out-of-range access"
        calls.append(rc)

    monkeypatch.setattr(mg.ops, "call", call)
    monkeypatch.setattr(mg.ops, "_need_cuda", lambda *ts: next(t.device for t in ts if t is not None))
    monkeypatch.setattr(mg.ops, "_stream", lambda: 0)
    monkeypatch.setattr(torch.cuda, "device", _NullCtx)
    return mg, calls


def test_module_and_wrapper_end_to_end_on_the_emulated_kernel(routed, golden):
    mg, calls = routed
    g = golden("fusion.npz")
    T = lambda a: torch.from_numpy(np.asarray(a))          # noqa: E731
    for tag in ["rand_i64", "rand_i32", "invalid", "blocky", "odd", "allbad"]:
        fu, table, m, ref = T(g[f"rg_{tag}_fu"]), T(g[f"rg_{tag}_table"]), T(g[f"rg_{tag}_map"]), T(g[f"rg_{tag}_out"])
        mod = mg.FeatureFusion([fu.shape[1]], table.shape[1])
        out = mod([fu], table, region_to_pixel_map=m)
        assert out.dtype == torch.float32 and torch.equal(out, ref), tag
        # bf16 fused buffer supplied by the caller
        buf = torch.empty(ref.shape, dtype=torch.bfloat16)
        out16 = mod([fu.bfloat16()], table, region_to_pixel_map=m, out=buf)
        assert out16.data_ptr() == buf.data_ptr()
        assert torch.equal(out16[:, fu.shape[1]:], ref[:, fu.shape[1]:].bfloat16()) and torch.equal(out16[:, :fu.shape[1]], fu.bfloat16())
    assert len(calls) == 12 and calls.count(1) >= 6               # the vector kernel served the aligned cases
    # uint8 map (the reference calls .long() on it), add fusion, dense branches
    out = mg.FeatureFusion([16], 16, "add")([T(g["add_fu"])], T(g["add_table"]), region_to_pixel_map=T(g["add_map"]))
    assert torch.equal(out, T(g["add_out"]))
    m8 = torch.randint(0, 5, (2, 8, 8), dtype=torch.uint8)
    out = mg.FeatureFusion([16], 16)([T(g["add_fu"])], T(g["add_table"]), region_to_pixel_map=m8)
    assert torch.equal(out, O.feature_fusion([T(g["add_fu"])], T(g["add_table"]), region_to_pixel_map=m8))
    out = mg.FeatureFusion([6], 10)([T(g["d4_same_fu"])], T(g["d4_same_fg"]), target_spatial_size=(16, 16))
    assert torch.equal(out, T(g["d4_same_out"]))
    out = mg.FeatureFusion([4, 3], 8)([T(g["d4_resize_fu0"]), T(g["d4_resize_fu1"])], T(g["d4_resize_fg"]))
    assert torch.allclose(out, T(g["d4_resize_out"]), atol=1e-6, rtol=0)
    # inputs that already are the channel slices of `out` are not copied (scope row f1)
    fused = torch.randn(2, 32 + 64, 16, 8)
    keep = fused.clone()
    out = mg.FeatureFusion([32], 64)([fused[:, :32]], fused[:, 32:], out=fused)
    assert out.data_ptr() == fused.data_ptr() and torch.equal(out, keep)
    # wrapper straight into a channel slice of a wider buffer; neighbours untouched
    table = torch.randn(9, 64)
    m = torch.randint(-1, 10, (3, 16, 32))
    wide = torch.full((3, 32 + 64 + 4, 16, 32), 7.0)
    mg.ops.region_map_gather(table, m, out=wide[:, 32:96])
    assert torch.equal(wide[:, 32:96], O.region_map_gather(table, m))
    assert bool((wide[:, :32] == 7.0).all()) and bool((wide[:, 96:] == 7.0).all())
    with pytest.raises(ValueError):
        mg.ops.region_map_gather(table, m, out=torch.empty(3, 64, 16, 33))
    with pytest.raises(IndexError):
        mg.FeatureFusion([32], 64)([torch.zeros(3, 32, 16, 32)], table, region_to_pixel_map=m[:, :8])
    with pytest.raises(NotImplementedError):
        mg.FeatureFusion([32], 64)([torch.zeros(3, 32, 16, 32)], table.clone().requires_grad_(True), region_to_pixel_map=m)
