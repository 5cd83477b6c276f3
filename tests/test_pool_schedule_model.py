"""CPU: a step-by-step model of pool_patches_tma_kernel's per-warp schedule (csrc/pool_unpool.cu) — issue cursor,
consume cursor, ring stage / mbarrier phase arithmetic, and the strip FIFO of the dynamic variant — run with random
interleavings of the warps.  It checks the invariants the kernel relies on: every (strip, chunk) is copied exactly once,
a warp consumes chunks in the order it issued them and always finds the ring stage in the phase it waits for, a stage
is never re-armed before it was drained, and the FIFO neither overflows nor is read before it was written.
(The GPU parity tests check the kernel's results; this pins the scheduling argument, in particular for the
MG_POOL_DYNAMIC variant written after the round-1 GPU budget was spent.)"""
import random

import pytest

K_WARPS, K_FIFO = 8, 16


class Cursor:
    def __init__(self):
        self.s = self.r0 = self.rows = 0

    def load(self, shape):
        nstrips, C, Hp, ph, Hf = shape
        if self.s >= nstrips:
            return
        py = (self.s // C) % Hp
        y0 = py * ph
        self.rows = min(Hf, y0 + ph) - y0
        self.r0 = 0


class Warp:
    """One consumer/producer warp, advanced one kernel-level action at a time."""

    def __init__(self, cta, warp, grid, shape, q, rpc, dynamic, counter, log):
        self.shape, self.q, self.rpc, self.dynamic, self.counter, self.log = shape, q, rpc, dynamic, counter, log
        self.nstrips = shape[0]
        self.fifo, self.tail, self.head = [None] * K_FIFO, 0, 0
        self.step = K_WARPS * grid
        self.stage_fill = [0] * q          # completed fills per ring stage (= mbarrier phases completed)
        self.stage_busy = [False] * q      # armed and not yet drained
        self.pending = []                  # chunks issued, in order: (strip, r0, stage, fill index)
        self.issued = self.consumed = 0
        self.ic, self.cc = Cursor(), Cursor()
        if dynamic:
            self.ic.s = self.draw()
            self.cc.s = self.follow()
        else:
            self.ic.s = self.cc.s = cta + warp * grid
        self.ic.load(shape)
        self.cc.load(shape)
        for _ in range(q):
            self.issue()
        self.done = not self.cc.s < self.nstrips

    def draw(self):
        s = self.counter[0]
        self.counter[0] += 1
        assert self.tail - self.head < K_FIFO, "FIFO overflow"
        self.fifo[self.tail % K_FIFO] = s
        self.tail += 1
        return s

    def follow(self):
        assert self.head < self.tail, "FIFO read before it was written"
        s = self.fifo[self.head % K_FIFO]
        self.head += 1
        return s

    def issue(self):
        if not self.ic.s < self.nstrips:
            return
        st = self.issued % self.q
        assert not self.stage_busy[st], "stage re-armed before it was drained"
        self.stage_busy[st] = True
        self.pending.append((self.ic.s, self.ic.r0, st, self.issued // self.q))
        self.log.append((self.ic.s, self.ic.r0))
        self.issued += 1
        self.ic.r0 += self.rpc
        if self.ic.r0 >= self.ic.rows:
            self.ic.s = self.draw() if self.dynamic else self.ic.s + self.step
            self.ic.load(self.shape)

    def consume_one(self):
        st = self.consumed % self.q
        strip, r0, ist, fill = self.pending.pop(0)
        # the chunk at the head of the issue order is the one the consume cursor expects, in the stage and phase it waits on
        assert (strip, r0, ist, fill) == (self.cc.s, self.cc.r0, st, self.consumed // self.q)
        assert self.stage_busy[st] and self.stage_fill[st] == fill
        self.stage_fill[st] += 1
        self.stage_busy[st] = False                      # __syncwarp(): drained
        self.issue()
        self.consumed += 1
        self.cc.r0 += self.rpc
        if self.cc.r0 >= self.cc.rows:
            self.cc.s = self.follow() if self.dynamic else self.cc.s + self.step
            self.cc.load(self.shape)
        self.done = not self.cc.s < self.nstrips


@pytest.mark.parametrize("dynamic", [False, True])
@pytest.mark.parametrize("B,C,Hf,ph,rpc,q,grid", [
    (2, 20, 512, 16, 16, 3, 148),      # cfg 2: one chunk per strip
    (1, 3, 70, 16, 4, 2, 5),           # ragged bottom strip, 4 chunks per strip, more warps than strips on some CTAs
    (3, 7, 100, 32, 8, 8, 4),          # deep ring
    (1, 1, 16, 16, 1, 2, 1),           # 16 chunks for a single strip
    (2, 5, 33, 8, 3, 3, 7),
])
def test_pool_schedule(B, C, Hf, ph, rpc, q, grid, dynamic):
    Hp = -(-Hf // ph)
    nstrips = B * Hp * C
    shape = (nstrips, C, Hp, ph, Hf)
    rng = random.Random(1234 + nstrips + q)
    counter, log = [0], []
    order = [(c, w) for c in range(grid) for w in range(K_WARPS)]
    rng.shuffle(order)                                    # CTAs / warps become resident in any order
    warps = []
    for i, (c, w) in enumerate(order):
        # late arrivals: interleave construction (prologue issues) with other warps' progress
        warps.append(Warp(c, w, grid, shape, q, rpc, dynamic, counter, log))
        for _ in range(rng.randrange(0, 4)):
            live = [x for x in warps if not x.done]
            if live:
                rng.choice(live).consume_one()
    while True:
        live = [x for x in warps if not x.done]
        if not live:
            break
        rng.choice(live).consume_one()
    want = []
    for s in range(nstrips):
        py = (s // C) % Hp
        rows = min(Hf, py * ph + ph) - py * ph
        want += [(s, r0) for r0 in range(0, rows, rpc)]
    assert sorted(log) == want                            # every chunk of every strip exactly once
    assert all(not x.pending and x.issued == x.consumed for x in warps)
    if dynamic:
        assert counter[0] == nstrips + grid * K_WARPS     # each warp draws exactly one terminating id


def test_mma_pool_fragment_mapping_sums_every_patch_row_once():
    """pool_patches_mma_kernel (csrc/pool_unpool.cu, opt-in, not yet run on hardware): model of its ldmatrix lane
    addresses and of the m16n8k16 fragment layouts.  512 contiguous bytes of an image row = 16 patches x 16 bf16 pixels;
    with B = ones the MMA must add, into D row i, the 16 pixels of patch i — each exactly once — and the 8 row addresses
    of every ldmatrix phase must fall into 8 different 16-byte bank groups (rows are 32 bytes apart)."""
    import numpy as np
    rng = np.random.default_rng(0)
    tile = rng.integers(-8, 9, size=(16, 16)).astype(np.float64)          # [patch][pixel], row-major = the smem bytes
    flat = tile.reshape(-1)                                                # element index = byte offset / 2

    def lane_off(lane):                                                    # bytes, as in the kernel
        return ((lane & 7) + ((lane >> 3) & 1) * 8) * 32 + ((((lane >> 4) & 1) ^ ((lane >> 2) & 1)) * 16)

    # ldmatrix.x4: matrix m, row j comes from the address of lane 8 m + j (16 bytes = 8 elements)
    mats = np.zeros((4, 8, 8))
    for m in range(4):
        offs = [lane_off(8 * m + j) for j in range(8)]
        assert len({(o % 128) // 16 for o in offs}) == 8                   # conflict-free phase
        for j, o in enumerate(offs):
            assert o % 16 == 0
            mats[m, j] = flat[o // 2:o // 2 + 8]
    # mma A fragment: reg0 = matrix 0 -> A[g][k 0..7], reg1 = matrix 1 -> A[g+8][k 0..7], reg2 -> A[g][8..15], reg3 -> A[g+8][8..15]
    Am = np.zeros((16, 16))
    Am[0:8, 0:8], Am[8:16, 0:8], Am[0:8, 8:16], Am[8:16, 8:16] = mats[0], mats[1], mats[2], mats[3]
    D = Am @ np.ones((16, 8))
    for i in range(16):
        assert sorted(Am[i]) == sorted(tile[i])                            # A row i = the 16 pixels of patch i, once each
    # D fragment: c0 -> D[g][2t], c2 -> D[g+8][2t]; the kernel stores c0 / c2 of the lanes with t == 0 as patches g, g + 8
    for lane in range(0, 32, 4):
        g = lane >> 2
        assert D[g][0] == tile[g].sum() and D[g + 8][0] == tile[g + 8].sum()
