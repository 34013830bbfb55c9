"""Per-operator unit tests of the CPU oracle: exhaustive small-domain soundness, ground
completeness, entailment soundness and monotonicity (SURVEY.md §4 "Implication for the new build").
The oracle is the parity anchor of the CUDA kernels, so its own semantics are pinned here by
brute force against the mathematical relation  x = y op z.
"""
import itertools

import numpy as np
import pytest

from oracle import oracle_py as orc
from turbo_b200 import abi

R = range(-4, 5)


def tdiv(a, b):
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b > 0) else -q


def tmod(a, b):
    return a - b * tdiv(a, b)


REL = {
    abi.OP_ADD: lambda x, y, z: x == y + z,
    abi.OP_MUL: lambda x, y, z: x == y * z,
    abi.OP_TDIV: lambda x, y, z: z != 0 and x == tdiv(y, z),
    abi.OP_TMOD: lambda x, y, z: z != 0 and x == tmod(y, z),
    abi.OP_MIN: lambda x, y, z: x == min(y, z),
    abi.OP_MAX: lambda x, y, z: x == max(y, z),
    abi.OP_EQ: lambda x, y, z: x in (0, 1) and x == int(y == z),
    abi.OP_LEQ: lambda x, y, z: x in (0, 1) and x == int(y <= z),
}


def boxes(rng, n, op):
    for _ in range(n):
        b = []
        for k in range(3):
            lo, hi = sorted(rng.integers(-4, 5, size=2))
            if k == 0 and op in (abi.OP_EQ, abi.OP_LEQ):
                lo, hi = sorted(rng.integers(0, 2, size=2))
            b.append((int(lo), int(hi)))
        yield b


def run_to_fixpoint(op, lb, ub):
    failed = False
    for _ in range(100):
        ch, f, lb, ub = orc.deduce(op, 0, 1, 2, lb, ub)
        failed |= f
        if not ch or failed:
            break
    return failed, lb, ub


@pytest.mark.parametrize("op", range(8))
def test_sound_on_small_boxes(op):
    rng = np.random.default_rng(op)
    for box in boxes(rng, 400, op):
        lb = [b[0] for b in box]
        ub = [b[1] for b in box]
        sols = [p for p in itertools.product(*[range(l, u + 1) for l, u in box]) if REL[op](*p)]
        failed, nlb, nub = run_to_fixpoint(op, lb, ub)
        if failed:
            assert not sols, (abi.OP_NAMES[op], box, sols[:3])
            continue
        for p in sols:
            for k in range(3):
                assert nlb[k] <= p[k] <= nub[k], (abi.OP_NAMES[op], box, p, nlb, nub)
        # never widens
        assert all(nlb[k] >= lb[k] and nub[k] <= ub[k] for k in range(3))


@pytest.mark.parametrize("op", range(8))
def test_ground_complete(op):
    """On fully assigned boxes deduce fails exactly when the relation is false, and ask agrees."""
    for x, y, z in itertools.product(R, R, R):
        if op in (abi.OP_EQ, abi.OP_LEQ) and x not in (0, 1):
            continue
        failed, nlb, nub = run_to_fixpoint(op, [x, y, z], [x, y, z])
        assert failed == (not REL[op](x, y, z)), (abi.OP_NAMES[op], x, y, z)
        if not failed:
            assert orc.ask(op, 0, 1, 2, nlb, nub)


@pytest.mark.parametrize("op", range(8))
def test_ask_is_sound(op):
    rng = np.random.default_rng(100 + op)
    seen = 0
    for box in boxes(rng, 600, op):
        # ask is only meaningful on a non-failed fixpoint (that is where propagate() calls it,
        # barebones_dive_and_solve.hpp:970-982)
        failed, lb, ub = run_to_fixpoint(op, [b[0] for b in box], [b[1] for b in box])
        if failed:
            continue
        if orc.ask(op, 0, 1, 2, lb, ub):
            seen += 1
            for p in itertools.product(*[range(l, u + 1) for l, u in zip(lb, ub)]):
                assert REL[op](*p), (abi.OP_NAMES[op], box, p)
    assert seen > 0


@pytest.mark.parametrize("op", range(8))
def test_monotone(op):
    """A ⊆ B  ⇒  deduce(A) ⊆ deduce(B): the condition that makes the fixpoint schedule-independent."""
    rng = np.random.default_rng(200 + op)
    for box in boxes(rng, 300, op):
        lbB = [b[0] for b in box]
        ubB = [b[1] for b in box]
        lbA, ubA = [], []
        for l, u in box:
            a, b = sorted(rng.integers(l, u + 1, size=2))
            lbA.append(int(a))
            ubA.append(int(b))
        fA, la, ua = run_to_fixpoint(op, lbA, ubA)
        fB, lB, uB = run_to_fixpoint(op, lbB, ubB)
        if fB:
            assert fA, (abi.OP_NAMES[op], box, lbA, ubA)
        if not fA and not fB:
            assert all(la[k] >= lB[k] and ua[k] <= uB[k] for k in range(3)), (abi.OP_NAMES[op], box, lbA, ubA)


def test_infinite_bounds():
    NI, PI = abi.NEG_INF, abi.POS_INF
    # x = y + z with x unbounded picks up finite bounds
    ch, f, lb, ub = orc.deduce(abi.OP_ADD, 0, 1, 2, [NI, 1, 2], [PI, 3, 4])
    assert ch and not f and (lb[0], ub[0]) == (3, 7)
    # an unbounded operand leaves the others alone
    ch, f, lb, ub = orc.deduce(abi.OP_ADD, 0, 1, 2, [0, NI, 2], [10, PI, 4])
    assert (lb[1], ub[1]) == (-4, 8)
    ch, f, lb, ub = orc.deduce(abi.OP_ADD, 0, 1, 2, [NI, NI, 2], [PI, PI, 4])
    assert not ch and not f
    # saturation instead of wrap-around
    big = 2 ** 31 - 2
    ch, f, lb, ub = orc.deduce(abi.OP_ADD, 0, 1, 2, [NI, big, big], [PI, big, big])
    assert lb[0] == PI and ub[0] == PI
    ch, f, lb, ub = orc.deduce(abi.OP_MUL, 0, 1, 2, [NI, 976000, 3000], [PI, 976000, 3000])
    assert lb[0] == PI  # 2.9e9 does not fit: saturates, never wraps negative
    # x <= y with unbounded sides
    ch, f, lb, ub = orc.deduce(abi.OP_LEQ, 0, 1, 2, [1, NI, NI], [1, PI, 5])
    assert ub[1] == 5 and lb[2] == NI
    ch, f, lb, ub = orc.deduce(abi.OP_LEQ, 0, 1, 2, [0, NI, NI], [0, PI, 5])
    assert not ch and lb[1] == NI and ub[2] == 5  # y > z, nothing finite to push
