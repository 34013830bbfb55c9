"""Oracle-level tests of the fixpoint and the dive-and-solve search (CPU only).

Pins the oracle's search semantics against brute-force enumeration on tiny networks, and checks
the two properties the GPU parity relies on: the fixpoint does not depend on the propagator
schedule, and the optimum does not depend on the EPS depth / strategy.
"""
import itertools

import numpy as np
import pytest

from oracle import oracle_py as orc
from turbo_b200 import abi
from tests import tnf_gen
from tests.test_oracle_ops import REL


def brute_force(pb):
    doms = [range(int(l), int(u) + 1) for l, u in zip(pb.lb, pb.ub)]
    best = None
    nsol = 0
    props = [(int(p["op"]), int(p["x"]), int(p["y"]), int(p["z"])) for p in pb.props]
    for a in itertools.product(*doms):
        if all(REL[op](a[x], a[y], a[z]) for op, x, y, z in props):
            nsol += 1
            if pb.obj_var >= 0 and (best is None or a[pb.obj_var] < best):
                best = a[pb.obj_var]
    return nsol, best


@pytest.mark.parametrize("seed", range(12))
def test_fixpoint_schedule_independent(seed):
    pb = tnf_gen.planted(60 + 13 * seed, 150 + 40 * seed, seed)
    ref = orc.fixpoint(pb)
    assert not ref["failed"]
    # the planted assignment survives propagation (soundness at network level)
    assert np.all(ref["lb"] <= pb.planted) and np.all(pb.planted <= ref["ub"])
    rng = np.random.default_rng(seed)
    for _ in range(4):
        order = rng.permutation(pb.nprops).astype(np.int32)
        r = orc.fixpoint(pb, order=order)
        assert not r["failed"]
        assert np.array_equal(r["lb"], ref["lb"]) and np.array_equal(r["ub"], ref["ub"])
    # idempotence
    again = orc.fixpoint(pb, ref["lb"], ref["ub"])
    assert again["sweeps"] == 1 and np.array_equal(again["lb"], ref["lb"])


@pytest.mark.parametrize("seed", range(40))
def test_solve_matches_bruteforce(seed):
    pb = tnf_gen.random_net(8, 7 + seed % 5, seed, lo=-3, hi=3)
    nsol, best = brute_force(pb)
    for depth in (0, 3):
        r = orc.solve(pb, depth=depth)
        assert r["exhaustive"]
        assert r["has_solution"] == (nsol > 0), (seed, depth)
        if nsol:
            assert r["objective"] == best, (seed, depth)
            # the reported point (lb of every variable) satisfies every propagator
            a = r["lb"]
            for p in pb.props:
                assert REL[int(p["op"])](int(a[p["x"]]), int(a[p["y"]]), int(a[p["z"]]))


@pytest.mark.parametrize("seed", range(10))
def test_strategy_and_depth_do_not_change_the_optimum(seed):
    results = set()
    for var_order in (abi.VAR_INPUT_ORDER, abi.VAR_FIRST_FAIL, abi.VAR_ANTI_FIRST_FAIL, abi.VAR_SMALLEST, abi.VAR_LARGEST):
        for val_order in (abi.VAL_MIN, abi.VAL_MAX, abi.VAL_SPLIT, abi.VAL_REVERSE_SPLIT):
            strat = [(var_order, val_order, list(range(3, 12))), (abi.VAR_FIRST_FAIL, abi.VAL_MIN, [])]
            pb = tnf_gen.random_net(14, 16, 1000 + seed, lo=-4, hi=4, strategies=strat)
            for depth in (0, 4):
                r = orc.solve(pb, depth=depth)
                assert r["exhaustive"]
                results.add((r["has_solution"], r["objective"]))
    assert len(results) == 1, results


def test_satisfaction_stops_at_first_solution():
    pb = tnf_gen.planted(40, 60, 7, objective=False)
    r = orc.solve(pb, depth=2)
    assert r["has_solution"] and not r["exhaustive"] and r["stats"]["solutions"] == 1
    for p in pb.props:
        assert REL[int(p["op"])](int(r["lb"][p["x"]]), int(r["lb"][p["y"]]), int(r["lb"][p["z"]]))


def test_parallel_oracle_agrees():
    for seed in range(6):
        pb = tnf_gen.random_net(16, 18, 2000 + seed, lo=-4, hi=4)
        a = orc.solve(pb, depth=5, nthreads=1)
        b = orc.solve(pb, depth=5, nthreads=4)
        assert (a["has_solution"], a["objective"], a["exhaustive"]) == (b["has_solution"], b["objective"], b["exhaustive"])


def test_dive_partitions_the_search_space():
    """Every solution of the root lies in exactly one dive subproblem (EPS soundness)."""
    pb = tnf_gen.random_net(8, 6, 77, lo=-2, hi=2, objective=False)
    depth = 3
    props = [(int(p["op"]), int(p["x"]), int(p["y"]), int(p["z"])) for p in pb.props]
    doms = [range(int(l), int(u) + 1) for l, u in zip(pb.lb, pb.ub)]
    sols = [a for a in itertools.product(*doms) if all(REL[op](a[x], a[y], a[z]) for op, x, y, z in props)]
    boxes = []
    idx = 0
    while idx < (1 << depth):
        d = orc.dive(pb, idx, depth)
        rem = d["remaining_depth"]
        if d["leaf_kind"] != 1:
            boxes.append((d["lb"], d["ub"]))
        idx = ((idx >> rem) + 1) << rem if d["leaf_kind"] else idx + 1
    for a in sols:
        n = sum(all(l[i] <= a[i] <= u[i] for i in range(len(a))) for l, u in boxes)
        assert n == 1, (a, n)
