"""CPU model of the in-kernel sum over ranks (dcreg_b200/csrc/peer_reduce.cuh), run under adversarial interleavings.

The device code cannot run here (no GPU), so - like tests/test_certificate_logic.py does for the gap certificate - the
protocol is restated step by step and a random scheduler interleaves the ranks' atomic steps:

  epoch e on rank r:  for every peer q: store the packet {value, epoch e} into pkt[q][e & 1][r]   (one atomic 8-byte store)
                      for every peer q: spin until pkt[r][e & 1][q] carries epoch e, then take its value
                      sum in rank order

Checked: no rank ever reads a slot that a faster peer has already overwritten with a later epoch (the two-slot parity
argument of the header comment), every rank obtains the same bits, the exchange never deadlocks, and a model with ONE
data slot does get corrupted under the same schedules (so the test can see the failure it guards against).
"""
import random

import numpy as np
import pytest


class Rank:
    def __init__(self, r, n, epochs, values):
        self.r, self.n, self.epochs, self.values = r, n, epochs, values
        self.e = 1
        self.phase = "post"
        self.todo = [q for q in range(n) if q != r]
        self.got = {}
        self.sums = []
        self.corrupt = False

    def done(self):
        return self.e > self.epochs


def run(n, epochs, slots, seed, greedy_rank=None):
    rng = random.Random(seed)
    vals = np.random.default_rng(seed).standard_normal((epochs + 1, n))
    data = [[[(0, 0.0)] * n for _ in range(slots)] for _ in range(n)]  # pkt[owner][slot][src] = (epoch, value)
    ranks = [Rank(r, n, epochs, vals) for r in range(n)]
    steps = 0
    while not all(k.done() for k in ranks):
        steps += 1
        assert steps < 200000, "deadlock"
        live = [k for k in ranks if not k.done()]
        k = rng.choice(live)
        if greedy_rank is not None and not ranks[greedy_rank].done() and rng.random() < 0.8:
            k = ranks[greedy_rank]                                       # one rank runs far ahead whenever it can
        e = k.e
        if k.phase == "post":
            q = k.todo.pop()
            data[q][e % slots][k.r] = (e, float(vals[e, k.r]))           # self-validating packet
            if not k.todo:
                k.phase, k.todo = "wait", [q2 for q2 in range(n) if q2 != k.r]
        else:
            q = k.todo[-1]
            ep, v = data[k.r][e % slots][q]
            if ep > e:                                                   # a later epoch already sits in the slot: e is lost
                k.corrupt = True
                ep = e
            if ep == e:
                k.got[q] = v
                k.todo.pop()
                if not k.todo:
                    k.got[k.r] = float(vals[e, k.r])
                    s = 0.0
                    for r in range(n):
                        s += k.got[r]
                    k.sums.append(s)
                    k.got = {}
                    k.e += 1
                    k.phase, k.todo = "post", [q2 for q2 in range(n) if q2 != k.r]
            # else: spin (no state change)
    return ranks, vals


@pytest.mark.parametrize("n", [2, 3, 8])
def test_two_slot_mailboxes_never_lose_a_contribution(n):
    for seed in range(40):
        ranks, vals = run(n, epochs=25, slots=2, seed=seed, greedy_rank=seed % n if seed % 3 else None)
        assert not any(k.corrupt for k in ranks)
        for e in range(1, 26):
            ref = 0.0
            for r in range(n):
                ref += float(vals[e, r])
            for k in ranks:
                assert k.sums[e - 1] == ref                              # same order => same bits on every rank


def test_single_slot_model_is_caught():
    """With ONE data slot a fast rank overwrites epoch e before a slow peer has read it: the model must notice."""
    bad = 0
    for seed in range(60):
        ranks, _ = run(3, epochs=25, slots=1, seed=seed, greedy_rank=seed % 3)
        bad += any(k.corrupt for k in ranks)
    assert bad > 0
