"""CPU model of the peer-memory exchange protocol (csrc/peer_push.cu + distributed.PeerGather), which was written
without access to a GPU: R ranks, ``depth`` slots, random interleavings of the individual memory actions.

Per rank and step k (slot s = k % depth):
  push:  for every peer p, in any order and interleaved with everybody else's actions:
            payload[p][s][rank] = k                       (the 16-byte stores)
            flag[p][s][rank]    = ++seq[rank][s][p]       (st.release.sys after the fence + barrier)
  wait:  need[r] = ++wseq[rank][s][r]; passes once flag[rank][s][r] >= need[r] for every r (ld.acquire.sys)
  read:  payload[rank][s][:]  must be k for every source rank.
Flow control as in tools/check_captured_gather.py --mode p2p: a barrier every ``depth`` steps, so a slot is never
overwritten before its consumers read it.

The second test shows why ``PeerGather.reset()`` exists: if the producers' counters start ahead (the graphs' warm-up
passes push too), the first wait of a slot is satisfied by a warm-up push and a consumer can read a stale payload."""
import random

import pytest


def run(R, depth, steps, seed, warm_pushes=0):
    rng = random.Random(seed)
    payload = [[[-1] * R for _ in range(depth)] for _ in range(R)]       # [owner][slot][source]
    flag = [[[0] * R for _ in range(depth)] for _ in range(R)]
    seq = [[[0] * R for _ in range(depth)] for _ in range(R)]            # [producer][slot][peer]
    wseq = [[[0] * R for _ in range(depth)] for _ in range(R)]           # [consumer][slot][source]
    for r in range(R):                                                   # warm-up pushes (payload -1 = garbage)
        for s in range(depth):
            for p in range(R):
                seq[r][s][p] = warm_pushes
                flag[p][s][r] = warm_pushes
    stale = []
    for k0 in range(0, steps, depth):                                    # between barriers: `depth` steps per rank
        # every rank's program for this window: a list of actions executed in program order PER (rank, peer) lane;
        # lanes of one push (one CTA per peer) and different ranks interleave freely
        lanes = []
        for r in range(R):
            prev_wait = None
            for k in range(k0, min(steps, k0 + depth)):
                s = k % depth
                for p in range(R):
                    lanes.append([("store", r, p, s, k), ("flag", r, p, s, k)])
                # the consumer side of rank r for step k runs on the slot's stream after its own push kernel; model it
                # as a lane that may start any time (it only depends on flags)
                lanes.append([("wait", r, s, k)])
        need = {}
        for r in range(R):
            for k in range(k0, min(steps, k0 + depth)):
                s = k % depth
                for src in range(R):
                    wseq[r][s][src] += 1
                    need[(r, s, src)] = wseq[r][s][src]
        pending = [l for l in lanes]
        guard = 0
        while pending:
            guard += 1
            assert guard < 10 ** 6, "protocol model does not terminate"
            lane = rng.choice(pending)
            act = lane[0]
            if act[0] == "store":
                _, r, p, s, k = act
                payload[p][s][r] = k
            elif act[0] == "flag":
                _, r, p, s, k = act
                seq[r][s][p] += 1
                flag[p][s][r] = seq[r][s][p]
            else:
                _, r, s, k = act
                if not all(flag[r][s][src] >= need[(r, s, src)] for src in range(R)):
                    continue                                             # still spinning: pick another action
                if any(payload[r][s][src] != k for src in range(R)):
                    stale.append((r, k, list(payload[r][s])))
            lane.pop(0)
            if not lane:
                pending.remove(lane)
    return stale


@pytest.mark.parametrize("R,depth", [(2, 2), (4, 3), (8, 3)])
def test_wait_always_sees_the_step_payload_after_reset(R, depth):
    for seed in range(30):
        assert run(R, depth, steps=4 * depth + 1, seed=seed) == []


def test_without_reset_a_wait_can_pass_on_a_warmup_push():
    assert any(run(4, 3, steps=6, seed=seed, warm_pushes=2) for seed in range(30))
