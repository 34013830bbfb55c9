"""Multi-GPU tests (need >= 2 devices; skipped on a 1-GPU box): subproblem sharding over GPUs and the
peer-mapped incumbent cells."""
import threading

import numpy as np
import pytest

from tests import golden_io, tnf_gen
from turbo_b200 import abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from turbo_b200 import engine
    if engine.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    return engine


def run_all(solvers):
    res = [None] * len(solvers)

    def run(g):
        res[g] = solvers[g].solve()
    th = [threading.Thread(target=run, args=(g,)) for g in range(len(solvers))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return res


def solve_on(eng, pb, ngpu, **kw):
    solvers = [eng.Solver(pb, device=g, gpu_rank=g, gpu_world=ngpu, **kw) for g in range(ngpu)]
    eng.link_peers(solvers)
    res = run_all(solvers)
    bounds = [s.read_bound() for s in solvers]
    for s in solvers:
        s.close()
    return res, bounds


def test_sharded_search_finds_the_same_optimum(eng):
    from oracle import oracle_py as orc
    n = min(eng.device_count(), 8)
    for seed in range(10):
        pb = tnf_gen.search_instance(seed)
        o = orc.solve(pb, depth=0)
        res, bounds = solve_on(eng, pb, n, subproblems_power=6)
        objs = [r["objective"] for r in res if r["has_solution"]]
        assert all(r["exhaustive"] for r in res)
        assert (min(objs) if objs else None) == o["objective"], seed
        if objs:
            # every GPU's incumbent cell converged to the global optimum through the peer writes
            assert all(b == o["objective"] for b in bounds), (seed, bounds)
        # every subproblem is accounted for exactly once over the GPUs, whoever solved it (its owner or a thief)
        total = sum(r["stats"]["eps_solved_subproblems"] + r["stats"]["eps_skipped_subproblems"] for r in res)
        assert total == 64, (seed, total)


def test_accap_a3_bound_sharing(eng):
    """BASELINE config 3 (accap_a3, incumbent sharing across GPUs), bounded by a timeout."""
    pb, info = golden_io.load("accap_a3")
    n = min(eng.device_count(), 8)
    res, bounds = solve_on(eng, pb, n, timeout_ms=5000)
    objs = [r["objective"] for r in res if r["has_solution"]]
    assert objs
    best = min(objs)
    assert min(bounds) == best
    # all cells agree (the writer pushes to every peer)
    assert len(set(bounds)) == 1, bounds


def test_repeated_runs_of_linked_solvers_are_independent(eng):
    """the incumbent of a run must not leak into the next run of the same linked solvers (a stale bound
    prunes everything at or above it: has_solution = 0 with exhaustive = 1, a false UNSAT). The cells carry the run's
    epoch, so nothing has to be reset between the runs."""
    from oracle import oracle_py as orc
    n = min(eng.device_count(), 8)
    pb = tnf_gen.search_instance(3)
    o = orc.solve(pb, depth=0)
    assert o["has_solution"]
    solvers = [eng.Solver(pb, device=g, gpu_rank=g, gpu_world=n, subproblems_power=6) for g in range(n)]
    eng.link_peers(solvers)
    for run in range(4):
        res = run_all(solvers)
        objs = [r["objective"] for r in res if r["has_solution"]]
        assert all(r["exhaustive"] for r in res)
        assert objs and min(objs) == o["objective"], (run, objs)
        assert all(s.read_bound() == o["objective"] for s in solvers), run
    for s in solvers:
        s.close()


def test_final_gather_over_the_gpus(eng):
    """tb_result_pack on every GPU + tb_result_reduce = reduce_blocks across GPUs (what bench.py gathers with NCCL)."""
    from oracle import oracle_py as orc
    n = min(eng.device_count(), 8)
    for seed in range(6):
        pb = tnf_gen.search_instance(seed)
        o = orc.solve(pb, depth=0)
        solvers = [eng.Solver(pb, device=g, gpu_rank=g, gpu_world=n, subproblems_power=5) for g in range(n)]
        eng.link_peers(solvers)
        res = run_all(solvers)
        m = eng.result_reduce([s.result_pack() for s in solvers])
        for s in solvers:
            s.close()
        assert m["has_solution"] == o["has_solution"] and m["exhaustive"]
        if o["has_solution"]:
            assert int(m["lb"][pb.obj_var]) == o["objective"]
            assert res[m["best_rank"]]["objective"] == o["objective"]
        assert m["stats"]["nodes"] == sum(r["stats"]["nodes"] for r in res)
        assert m["stats"]["eps_solved_subproblems"] + m["stats"]["eps_skipped_subproblems"] == 32


def test_an_idle_gpu_steals_from_its_peers(eng):
    """GPU 0 searches with a single block, so GPU 1 runs out of subproblems of its own shard first and takes the
    slow GPU's through the peer-mapped dispenser; every subproblem is still handed out at most once."""
    pb, info = golden_io.load_simplified_problem("accap_a3")
    solvers = [eng.Solver(pb, device=0, gpu_rank=0, gpu_world=2, subproblems_power=12, or_blocks=1, timeout_ms=3000),
               eng.Solver(pb, device=1, gpu_rank=1, gpu_world=2, subproblems_power=12, timeout_ms=3000)]
    eng.link_peers(solvers)
    res = run_all(solvers)
    for s in solvers:
        s.close()
    assert any(r["has_solution"] for r in res)
    assert res[1]["stats"]["eps_stolen_subproblems"] > 0 and res[0]["stats"]["eps_stolen_subproblems"] == 0
    assert sum(r["stats"]["eps_solved_subproblems"] + r["stats"]["eps_skipped_subproblems"] for r in res) <= 4096
    # and with static shards (TB_STEAL=0 at link time) nothing is stolen
    import os
    os.environ["TB_STEAL"] = "0"
    try:
        solvers = [eng.Solver(pb, device=0, gpu_rank=0, gpu_world=2, subproblems_power=12, or_blocks=1, timeout_ms=1000),
                   eng.Solver(pb, device=1, gpu_rank=1, gpu_world=2, subproblems_power=12, timeout_ms=1000)]
        eng.link_peers(solvers)
        res = run_all(solvers)
        for s in solvers:
            s.close()
    finally:
        del os.environ["TB_STEAL"]
    assert all(r["stats"]["eps_stolen_subproblems"] == 0 for r in res)


def test_the_tail_is_shared_between_the_gpus(eng, monkeypatch):
    """One subproblem for two GPUs: rank 1 has nothing of its own, its blocks wait, rank 0's busy block splits its
    subproblem and rank 1 takes children out of rank 0's pool (peer-mapped with the cell block). The search is still
    exhaustive and ends on the reference's optimum."""
    monkeypatch.setenv("TB_SPLIT_MIN_NODES", "256")
    for name in ("pat13", "triangular9"):
        pb, info = golden_io.load(name)
        solvers = [eng.Solver(pb, device=g, gpu_rank=g, gpu_world=2, subproblems_power=0, timeout_ms=60000) for g in range(2)]
        eng.link_peers(solvers)
        res = run_all(solvers)
        m = eng.result_reduce([s.result_pack() for s in solvers])
        for s in solvers:
            s.close()
        assert m["has_solution"] and m["exhaustive"], name
        assert golden_io.user_objective(info, m["lb"], m["ub"]) == info["expected"], name
        assert res[0]["stats"]["eps_split_subproblems"] > 0
        assert res[1]["stats"]["eps_split_parts_solved"] > 0 and res[1]["stats"]["nodes"] > 0      # children of rank 0's subproblem
    # TB_SHARE_SPLIT=0 at link time: every GPU keeps its tail to itself, same answer
    monkeypatch.setenv("TB_SHARE_SPLIT", "0")
    pb, info = golden_io.load("pat13")
    solvers = [eng.Solver(pb, device=g, gpu_rank=g, gpu_world=2, subproblems_power=0, timeout_ms=60000) for g in range(2)]
    eng.link_peers(solvers)
    res = run_all(solvers)
    m = eng.result_reduce([s.result_pack() for s in solvers])
    for s in solvers:
        s.close()
    assert m["exhaustive"] and golden_io.user_objective(info, m["lb"], m["ub"]) == info["expected"]
    assert res[1]["stats"]["eps_split_parts_solved"] == 0


def test_first_solution_of_a_satisfaction_problem_stops_every_gpu(eng):
    n = min(eng.device_count(), 8)
    pb = tnf_gen.planted(60, 80, 5, objective=False)
    solvers = [eng.Solver(pb, device=g, gpu_rank=g, gpu_world=n, subproblems_power=10) for g in range(n)]
    eng.link_peers(solvers)
    res = run_all(solvers)
    for s in solvers:
        s.close()
    assert any(r["has_solution"] for r in res)
    # the GPUs that found nothing were stopped by the one that did: none of them claims to have exhausted its shard
    # unless it really had nothing left
    assert sum(r["stats"]["nodes"] for r in res) < 1024 * 50
