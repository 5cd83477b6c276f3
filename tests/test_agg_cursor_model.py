"""CPU: a model of the work cursor of the wide-row aggregation kernel (csrc/gat_tc.cu, gat_agg_spill_kernel) and of the
three-deep software pipeline that consumes it.

A warp owns the steps ``gw, gw + nw, gw + 2 nw, ...`` (a step = 4 consecutive destinations).  Per step it walks the
``nsl`` slabs of 64 features and, per slab, the ``nch`` chunks of 8 in-edges (``nch`` = the largest in-degree of the
step's 4 destinations, rounded up to chunks, at least 1): sub-item = (step, slab, chunk).  The cursor is advanced TWO
sub-items ahead of the one on the tensor cores (column ids of n+2 and rows / attention scalars of n+1 are in flight
while n computes), row pointers of the step after the cursor's are prefetched in place, and with one chunk per step
the column ids, attention scalars and numerator fragments of a step are loaded / computed once and reused for the
following slabs (``reuse``).  The properties checked here are the ones the kernel relies on:

* every (step, slab, chunk) of every warp is produced exactly once, slabs and chunks in order, ``first`` / ``last`` set on
  the first / last chunk of a (step, slab) only (the accumulators are zeroed / the z slab is stored exactly once);
* a reused sub-item follows a sub-item of the same step and chunk whose values it may take over, and only when the
  step has a single chunk;
* the prefetched row pointers are those of the step they are later used for, including past the end of the graph
  (an empty range, read from ``rowptr[N]``);
* the edge slots covered by all sub-items of a step are exactly the in-edges of its 4 destinations, each once.
"""
import random

import pytest


class Cursor:
    """Lane-level replica of the cursor in gat_agg_spill_kernel for loader lane group r4 (= destination of the step)."""

    def __init__(self, rowptr, N, nsl, gw, nw):
        self.rowptr, self.N, self.nsl, self.nw = rowptr, N, nsl, nw
        self.nsteps = (N + 3) // 4
        self.step, self.sl, self.c = gw, 0, 0
        self.beg, self.end = [0] * 4, [0] * 4
        self.nbeg, self.nend = [0] * 4, [0] * 4
        self.prefetched_for = None
        self._fetch(self.step, self.beg, self.end)
        self._fetch(self.step + nw, self.nbeg, self.nend)
        self.prefetched_for = self.step + nw
        self.nch = max(1, max((self.end[r] - self.beg[r] + 7) >> 3 for r in range(4)))

    def _fetch(self, step, b, e):
        for r4 in range(4):
            j = min(step * 4 + r4, self.N) if step < self.nsteps else self.N
            b[r4], e[r4] = self.rowptr[j], self.rowptr[min(j + 1, self.N)]

    def advance(self):
        adv_step = False
        self.c += 1
        if self.c >= self.nch:
            self.c = 0
            self.sl += 1
            if self.sl >= self.nsl:
                self.sl = 0
                self.step += self.nw
                adv_step = True
        if adv_step:
            assert self.prefetched_for == self.step          # the in-place prefetch was for exactly this step
            self.beg, self.end = list(self.nbeg), list(self.nend)
            self._fetch(self.step + self.nw, self.nbeg, self.nend)
            self.prefetched_for = self.step + self.nw
            self.nch = max(1, max((self.end[r] - self.beg[r] + 7) >> 3 for r in range(4)))

    def item(self):
        node0 = self.step * 4 if self.step < self.nsteps else -1
        return dict(node0=node0, step=self.step, sl=self.sl, c=self.c, first=self.c == 0, last=self.c == self.nch - 1,
                    reuse=self.nch == 1 and self.sl > 0, nch=self.nch)

    def slots(self):
        """edge ids of the 32 (destination r4, slot ch8) loader lanes of the cursor's sub-item (-1: empty)."""
        out = []
        for r4 in range(4):
            for ch8 in range(8):
                k = self.beg[r4] + 8 * self.c + ch8
                out.append(k if (self.step < self.nsteps and k < self.end[r4]) else -1)
        return out


def run_warp(rowptr, N, nsl, gw, nw):
    """The kernel's pipeline for one warp: returns the sub-items in the order they reach the tensor cores, each with the
    edge slots whose rows were staged for it."""
    cur = Cursor(rowptr, N, nsl, gw, nw)
    it0, src0 = cur.item(), cur.slots()
    cur.advance()
    it1 = cur.item()
    src1 = src0 if it1["reuse"] else cur.slots()
    done = []
    while it0["node0"] >= 0:
        cur.advance()
        it2 = cur.item()
        src2 = src1 if it2["reuse"] else cur.slots()
        done.append((it0, src0))
        it0, src0, it1, src1 = it1, src1, it2, src2
    return done


def random_rowptr(rng, N, kmax):
    deg = [rng.choice([0, 1, kmax // 2, kmax, rng.randint(0, kmax)]) for _ in range(N)]
    rowptr = [0]
    for d in deg:
        rowptr.append(rowptr[-1] + d)
    return rowptr


@pytest.mark.parametrize("seed", range(12))
def test_cursor_enumerates_every_sub_item_once(seed):
    rng = random.Random(seed)
    N = rng.choice([1, 3, 4, 5, 37, 64, 203])
    nsl = rng.choice([1, 2, 4, 8])
    kmax = rng.choice([3, 8, 9, 20])
    nw = rng.choice([1, 2, 5, 16])
    rowptr = random_rowptr(rng, N, kmax)
    nsteps = (N + 3) // 4
    seen_edges = set()
    seen_items = set()
    for gw in range(nw):
        items = run_warp(rowptr, N, nsl, gw, nw)
        expect_step, prev = gw, None
        pos = 0
        while pos < len(items):
            it, _ = items[pos]
            assert it["step"] == expect_step and expect_step < nsteps
            nch = it["nch"]
            degs = [rowptr[min(it["node0"] + r + 1, N)] - rowptr[min(it["node0"] + r, N)] for r in range(4)]
            assert nch == max(1, max((d + 7) // 8 for d in degs))
            for sl in range(nsl):
                for c in range(nch):
                    cur_it, slots = items[pos]
                    assert (cur_it["step"], cur_it["sl"], cur_it["c"]) == (expect_step, sl, c)
                    assert cur_it["first"] == (c == 0) and cur_it["last"] == (c == nch - 1)
                    assert cur_it["reuse"] == (nch == 1 and sl > 0)
                    if cur_it["reuse"]:
                        assert prev is not None and prev[0]["step"] == expect_step and prev[0]["c"] == c and slots == prev[1]
                    key = (expect_step, sl, c)
                    assert key not in seen_items
                    seen_items.add(key)
                    # the staged slots are the chunk's in-edges of the 4 destinations
                    for r4 in range(4):
                        j = it["node0"] + r4
                        b, e = (rowptr[j], rowptr[j + 1]) if j < N else (rowptr[N], rowptr[N])
                        want = [k if k < e else -1 for k in range(b + 8 * c, b + 8 * c + 8)]
                        assert slots[r4 * 8:(r4 + 1) * 8] == want
                    if sl == 0:
                        for k in slots:
                            if k >= 0:
                                assert k not in seen_edges
                                seen_edges.add(k)
                    prev = items[pos]
                    pos += 1
            expect_step += nw
        assert expect_step >= nsteps                              # the warp stopped exactly at the end of its steps
    assert len(seen_items) == sum(nsl * max(1, max((rowptr[min(s * 4 + r + 1, N)] - rowptr[min(s * 4 + r, N)] + 7) // 8 for r in range(4)))
                                  for s in range(nsteps))
    assert seen_edges == set(range(rowptr[N]))                    # every in-edge of the graph exactly once (per slab)


def test_cursor_with_no_work_for_a_warp():
    """More warps than steps: a warp whose first step lies past the end produces nothing and reads only rowptr[N]."""
    rowptr = [0, 2, 2, 5]
    assert run_warp(rowptr, 3, 2, gw=1, nw=4) == []
    assert len(run_warp(rowptr, 3, 2, gw=0, nw=4)) == 2
