"""CPU: the peer-memory exchange kernels (csrc/peer_push.cu: peer_push_kernel / peer_wait_kernel, written after the
round-1 GPU budget was spent, not yet run on hardware) compiled for the host and executed thread by thread
(tests/emu/cuda_warp_shim.h), replaying what distributed.PeerGather does with them: ``world`` emulated GPUs, ``depth``
slots, several steps.  Checks the destination offsets, the flag indices, the device-side sequence counters and the
bounded wait."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
EMU = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.isfile(os.path.join(inc, "cuda_bf16.h")):
        pytest.skip("CUDA headers not available")
    so = str(tmp_path_factory.mktemp("emu") / "peer_emu.so")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-w", "-pthread", "-I", inc, "-o", so,
                        os.path.join(EMU, "peer_emu.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(so)
    lib.emu_peer_push.restype = C.c_int
    lib.emu_peer_push.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p]
    lib.emu_peer_wait.restype = C.c_int
    lib.emu_peer_wait.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_ulonglong]
    return lib


def aligned(n, dtype):
    raw = np.zeros(n * np.dtype(dtype).itemsize + 128, dtype=np.uint8)
    start = (-raw.ctypes.data) % 64
    return raw[start:start + n * np.dtype(dtype).itemsize].view(dtype)


@pytest.mark.parametrize("world,depth,n_pad", [(2, 2, 8), (3, 3, 1044)])
def test_push_and_wait_replay_of_peer_gather(emu, world, depth, n_pad):
    gathered = [aligned(depth * world * n_pad, np.float32) for _ in range(world)]      # one symmetric buffer per "GPU"
    flags = [aligned(depth * world, np.uint32) for _ in range(world)]
    seq = [np.zeros((depth, world), dtype=np.uint32) for _ in range(world)]            # producer counters of each rank
    wseq = [np.zeros((depth, world), dtype=np.uint32) for _ in range(world)]           # consumer counters of each rank
    status = np.zeros(1, dtype=np.int32)
    bufs = (C.c_void_p * world)(*[g.ctypes.data for g in gathered])
    sigs = (C.c_void_p * world)(*[f.ctypes.data for f in flags])
    payload = aligned(n_pad, np.float32)
    for k in range(2 * depth + 1):
        slot = k % depth
        for r in range(world):                                                         # every rank pushes its step-k payload
            payload[:] = 1000.0 * r + k + np.arange(n_pad) / 4096.0
            off = (slot * world + r) * n_pad * 4
            rc = emu.emu_peer_push(payload.ctypes.data, n_pad * 4, bufs, world, off, sigs, slot * world + r,
                                   seq[r][slot].ctypes.data)
            assert rc == 0
        for r in range(world):                                                         # every rank waits, then reads
            rc = emu.emu_peer_wait(flags[r].ctypes.data, slot * world, world, wseq[r][slot].ctypes.data, status.ctypes.data,
                                   1000)
            assert rc == 0 and status[0] == 0
            got = gathered[r].reshape(depth, world, n_pad)[slot]
            for src in range(world):
                assert np.array_equal(got[src], (1000.0 * src + k + np.arange(n_pad) / 4096.0).astype(np.float32))
            assert np.all(flags[r].reshape(depth, world)[slot] == k // depth + 1)
        assert all(np.all(seq[r][slot] == k // depth + 1) and np.all(wseq[r][slot] == k // depth + 1) for r in range(world))
    # other slots' data was not disturbed by the last step; a wait without a matching push expires and reports it
    rc = emu.emu_peer_wait(flags[0].ctypes.data, 0, world, wseq[0][0].ctypes.data, status.ctypes.data, 50)
    assert rc == 0 and status[0] == 1


def test_push_rejects_unaligned_payloads(emu):
    buf = aligned(64, np.float32)
    assert emu.emu_peer_push(buf.ctypes.data + 4, 16, None, 1, 0, None, 0, None) == -2
    assert emu.emu_peer_push(buf.ctypes.data, 24, None, 1, 0, None, 0, None) == -2
