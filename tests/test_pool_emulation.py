"""CPU: the device code of the bulk-copy staged pooling kernels (csrc/pool_tma_kernels.cuh) compiled for the HOST with g++
(tests/emu/cuda_warp_shim.h) and executed with every CUDA thread as a host thread: warp collectives as lock-step
exchanges, mbarriers as a state machine, bulk copies as bounds- and alignment-checked memcpys.  Covers the static strip
order (the GPU-tested default), the dynamic strip scheduling and the tensor-core summation variants (both written after
the round-1 GPU budget was spent and not yet run on hardware): results against a numpy window mean on integer-valued
inputs (every sum is then exact in fp32, whatever the order), every input byte copied exactly once, no access outside
the input / shared memory / output, counters of the dynamic variant re-armed."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
EMU = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.isfile(os.path.join(inc, "cuda_bf16.h")):
        pytest.skip("CUDA headers not available")
    so = str(tmp_path_factory.mktemp("emu") / "pool_emu.so")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-w", "-pthread", "-I", inc, "-o", so,
                        os.path.join(EMU, "pool_emu.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(so)
    lib.emu_pool.restype = C.c_int
    lib.emu_pool.argtypes = [C.c_void_p, C.c_int, C.c_longlong] + [C.c_int] * 6 + [C.c_void_p] + [C.c_int] * 5 + \
        [C.c_void_p, C.c_void_p]
    return lib


def run_pool(lib, B, Cc, Hf, Wf, ph, pw, x_bf16, out_bf16, variant, sms=2, stages=3, chunk=8192, seed=0, counters=None):
    gen = torch.Generator().manual_seed(seed)
    x = torch.randint(-4, 5, (B, Cc, Hf, Wf), generator=gen).float()          # small integers: exact in bf16, exact sums
    xs = x.bfloat16().contiguous() if x_bf16 else x.contiguous()
    Hp, Wp = -(-Hf // ph), -(-Wf // pw)
    odt = torch.bfloat16 if out_bf16 else torch.float32
    guard = 64
    buf = torch.full((guard + B * Hp * Wp * Cc + guard,), 777.0, dtype=odt)
    out = buf[guard:guard + B * Hp * Wp * Cc]
    copied = C.c_longlong(0)
    cnt = counters if counters is not None else (C.c_int * 2)(0, 0)
    rc = lib.emu_pool(xs.data_ptr(), int(x_bf16), xs.numel() * xs.element_size(), B, Cc, Hf, Wf, ph, pw, out.data_ptr(),
                      int(out_bf16), sms, variant, stages, chunk, C.addressof(cnt), C.addressof(copied))
    assert rc > 0, f"emulated kernel reported a fault or refused the shape (rc={rc})"
    # reference: zero padding counts in the mean (divisor ph * pw), layout (B, Hp*Wp, C)
    xp = torch.zeros(B, Cc, Hp * ph, Wp * pw)
    xp[:, :, :Hf, :Wf] = x
    ref = xp.reshape(B, Cc, Hp, ph, Wp, pw).sum(dim=(3, 5)) / float(ph * pw)
    ref = ref.permute(0, 2, 3, 1).reshape(B, Hp * Wp, Cc).to(odt)
    assert torch.equal(out.reshape(B, Hp * Wp, Cc), ref)
    assert bool((buf[:guard] == 777.0).all()) and bool((buf[-guard:] == 777.0).all())       # nothing written outside
    assert copied.value == xs.numel() * xs.element_size()                                   # every byte staged once
    if variant == 1:
        assert (cnt[0], cnt[1]) == (0, 0)                                                   # re-armed by the last CTA
    return rc


@pytest.mark.parametrize("x_bf16,out_bf16", [(True, True), (True, False), (False, False), (False, True)])
def test_static_strip_order_matches_window_mean(emu, x_bf16, out_bf16):
    # ragged last patch row (40 = 2 * 16 + 8), several strips per warp (2 CTAs x 8 warps for 2 * 3 * 3 = 18 strips)
    run_pool(emu, 2, 3, 40, 128, 16, 16, x_bf16, out_bf16, variant=0)
    # several chunks per strip (1 KB bf16 / 2 KB fp32 rows, 4 KB ring buffers), two stages
    run_pool(emu, 1, 2, 32, 512, 16, 16, x_bf16, out_bf16, variant=0, stages=2, chunk=4096, seed=1)


def test_other_patch_sizes(emu):
    run_pool(emu, 1, 5, 24, 64, 8, 8, True, True, variant=0, seed=2)        # one bf16 vector per patch column
    run_pool(emu, 2, 2, 12, 96, 4, 32, False, False, variant=0, seed=3)     # wide patches, fp32
    run_pool(emu, 1, 1, 16, 1024, 16, 16, True, False, variant=0, seed=4, sms=1)   # two column passes per lane, one CTA


@pytest.mark.parametrize("x_bf16", [True, False])
def test_dynamic_strip_scheduling(emu, x_bf16):
    cnt = (C.c_int * 2)(0, 0)
    for seed in range(2):                                                   # the same counter pair serves launch after launch
        run_pool(emu, 2, 3, 40, 128, 16, 16, x_bf16, x_bf16, variant=1, seed=seed, counters=cnt)
    run_pool(emu, 1, 2, 32, 512, 16, 16, x_bf16, False, variant=1, stages=2, chunk=4096, seed=5)
    run_pool(emu, 1, 1, 16, 128, 16, 16, x_bf16, False, variant=1, sms=3, seed=6)   # more warps than strips


@pytest.mark.parametrize("out_bf16", [True, False])
def test_tensor_core_summation_variant(emu, out_bf16):
    run_pool(emu, 2, 3, 40, 256, 16, 16, True, out_bf16, variant=2)                          # ragged last strip, one tile
    run_pool(emu, 1, 2, 32, 512, 16, 16, True, out_bf16, variant=2, stages=2, chunk=4096, seed=1)   # 2 tiles, 4 chunks per strip
    run_pool(emu, 1, 2, 24, 1024, 8, 16, True, out_bf16, variant=2, seed=2, sms=1)           # ph = 8, four tiles
