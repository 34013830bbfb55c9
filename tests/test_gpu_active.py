"""Active-set fixpoint (TB_FP_AC1_ACTIVE / TB_FP_WAC1_ACTIVE): a warp only evaluates the chunks of 32
propagators one of whose variables moved since the chunk was last evaluated.  The greatest fixpoint is unique, so every
store, every dive subproblem and the whole search trace (node, failure and solution counts) must be IDENTICAL to the
oracle's and to the dense sweeps'; only num_deductions may (and must) go down."""
import numpy as np
import pytest

from tests import golden_io, tnf_gen
from tests.test_gpu_parity import assert_same_store
from turbo_b200 import abi

pytestmark = pytest.mark.gpu

ACTIVE = [abi.FP_AC1_ACTIVE, abi.FP_WAC1_ACTIVE]
SHARED = [abi.MEM_TCN_SHARED, abi.MEM_STORE_SHARED]


@pytest.fixture(autouse=True)
def force_active_on_small_networks(monkeypatch):
    # the engine runs the plain sweeps below 128 chunks (they are faster there); these tests want the active path
    monkeypatch.setenv("TB_ACTIVE_MIN_CHUNKS", "0")


@pytest.fixture(scope="module")
def eng():
    from turbo_b200 import engine
    assert engine.device_count() > 0, "no CUDA device: the engine has no CPU fallback"
    return engine


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle_py
    return oracle_py


@pytest.mark.parametrize("fp", ACTIVE)
@pytest.mark.parametrize("kind", SHARED)
def test_root_fixpoint_bit_exact(eng, orc, kind, fp):
    for seed, (nv, npr) in enumerate([(8, 5), (33, 64), (100, 300), (1000, 3000), (5000, 20000), (20000, 6000), (27000, 60000)]):
        pb = tnf_gen.planted(nv, npr, seed)
        if kind == abi.MEM_TCN_SHARED and pb.nvars * 8 + pb.nprops * 8 > 190_000:
            continue
        o = orc.fixpoint(pb)
        with eng.Solver(pb, mem_kind=kind, fixpoint=fp) as s:
            assert s.config()["mem_kind"] == kind
            g = s.propagate()
            # a batch: the flags are reset between stores; perturbed stores start from another state
            rng = np.random.default_rng(seed)
            lb = np.tile(pb.lb, (3, 1)); ub = np.tile(pb.ub, (3, 1))
            pick = rng.integers(0, pb.nvars, size=5)
            for j in pick:
                if lb[1, j] < ub[1, j]:
                    lb[1, j] += 1
            if not o["failed"]:
                lb[2], ub[2] = o["lb"], o["ub"]
            b = s.propagate_batch(lb, ub)
        assert_same_store(g, o, (kind, fp, nv, npr))
        o1 = orc.fixpoint(pb, lb[1], ub[1])
        assert b["failed"][0] == o["failed"] and b["failed"][1] == o1["failed"]
        if not o1["failed"]:
            assert np.array_equal(b["lb"][1], o1["lb"]) and np.array_equal(b["ub"][1], o1["ub"])
        if not o["failed"]:
            assert np.array_equal(b["lb"][0], o["lb"]) and np.array_equal(b["lb"][2], o["lb"]) and np.array_equal(b["ub"][2], o["ub"])


@pytest.mark.parametrize("fp", ACTIVE)
def test_failed_stores_and_every_operator(eng, orc, fp):
    nfailed = 0
    for seed in range(40):
        pb = tnf_gen.random_net(20, 10 + seed, 9000 + seed, lo=-5, hi=5) if seed % 2 else tnf_gen.planted(40, 60 + seed, 9000 + seed)
        o = orc.fixpoint(pb)
        nfailed += o["failed"]
        for kind in SHARED:
            with eng.Solver(pb, mem_kind=kind, fixpoint=fp) as s:
                g = s.propagate()
            assert_same_store(g, o, (seed, kind, fp))
    assert 0 < nfailed < 40      # both regimes are covered


@pytest.mark.parametrize("name", golden_io.names())
def test_golden_root_fixpoints(eng, orc, name):
    pb, info = golden_io.load(name)
    o = orc.fixpoint(pb)
    with eng.Solver(pb, fixpoint=abi.FP_WAC1_ACTIVE) as s:
        g = s.propagate()
        d = None
        if s.config()["mem_kind"] in SHARED and pb.nprops > 64:
            with eng.Solver(pb, fixpoint=abi.FP_WAC1, mem_kind=s.config()["mem_kind"]) as s2:
                d = s2.propagate()
    assert_same_store(g, o, name)
    if d is not None and not o["failed"]:
        assert g["stats"]["num_deductions"] <= d["stats"]["num_deductions"], name


@pytest.mark.parametrize("fp", ACTIVE)
@pytest.mark.parametrize("kind", SHARED)
def test_every_dive_subproblem_bit_exact(eng, orc, kind, fp):
    for seed, depth in [(11, 4), (12, 6), (13, 5)]:
        strat = [(abi.VAR_INPUT_ORDER, abi.VAL_SPLIT, list(range(3, 40))), (abi.VAR_FIRST_FAIL, abi.VAL_MIN, [])]
        pb = tnf_gen.planted(120, 200, seed, strategies=strat, objective=True)
        with eng.Solver(pb, mem_kind=kind, fixpoint=fp) as s:
            g = s.dive_batch(0, 1 << depth, depth)
        for idx in range(1 << depth):
            o = orc.dive(pb, idx, depth)
            assert g["remaining_depth"][idx] == o["remaining_depth"], (seed, idx)
            assert g["leaf_kind"][idx] == o["leaf_kind"], (seed, idx)
            if o["leaf_kind"] != 1:
                assert np.array_equal(g["lb"][idx], o["lb"]) and np.array_equal(g["ub"][idx], o["ub"]), (seed, idx)


@pytest.mark.parametrize("fp", ACTIVE)
def test_single_block_search_trace_matches_oracle(eng, orc, fp):
    """Same fixpoint at every node => same tree: nodes, failures, solutions, subproblem counters and depth are equal."""
    for seed in range(12):
        pb = tnf_gen.search_instance(seed) if seed < 8 else tnf_gen.random_net(18, 10, 4000 + seed, lo=-4, hi=4)
        for depth in (0, 3):
            o = orc.solve(pb, depth=depth)
            for kind in SHARED:
                with eng.Solver(pb, or_blocks=1, subproblems_power=depth, fixpoint=fp, mem_kind=kind) as s:
                    g = s.solve()
                for key in ("nodes", "fails", "solutions", "eps_solved_subproblems", "eps_skipped_subproblems", "depth_max"):
                    assert g["stats"][key] == o["stats"][key], (seed, depth, kind, key, g["stats"][key], o["stats"][key])
                assert g["objective"] == o["objective"]


@pytest.mark.parametrize("fp", ACTIVE)
def test_solve_status_and_optimum_many_blocks(eng, orc, fp):
    for seed in range(30):
        pb = tnf_gen.search_instance(seed) if seed < 14 else tnf_gen.random_net(16, 9, 5000 + seed, lo=-4, hi=4)
        o = orc.solve(pb, depth=0)
        with eng.Solver(pb, subproblems_power=4, fixpoint=fp) as s:
            g = s.solve()
        assert g["exhaustive"] and o["exhaustive"]
        assert (g["has_solution"], g["objective"]) == (o["has_solution"], o["objective"]), seed


@pytest.mark.parametrize("name", [n for n in golden_io.names() if golden_io.load(n)[1]["expected"] is not None])
def test_golden_optimum_and_fewer_deductions(eng, name):
    pb, info = golden_io.load(name)
    with eng.Solver(pb, timeout_ms=120000, fixpoint=abi.FP_WAC1_ACTIVE) as s:
        r = s.solve()
    assert r["has_solution"] and r["exhaustive"]
    assert golden_io.user_objective(info, r["lb"], r["ub"]) == info["expected"]


def test_dense_and_active_traces_agree_on_a_real_network(eng):
    """One block, bounded node budget on the simplified trains15 network: the two fixpoints walk the same tree."""
    pb, info = golden_io.load_simplified_problem("trains15")
    res = {}
    for fp in (abi.FP_WAC1, abi.FP_WAC1_ACTIVE):
        with eng.Solver(pb, or_blocks=1, subproblems_power=6, cutnodes=400, fixpoint=fp) as s:
            res[fp] = s.solve()
    a, b = res[abi.FP_WAC1]["stats"], res[abi.FP_WAC1_ACTIVE]["stats"]
    for key in ("nodes", "fails", "solutions", "depth_max"):
        assert a[key] == b[key], (key, a[key], b[key])
    assert res[abi.FP_WAC1]["objective"] == res[abi.FP_WAC1_ACTIVE]["objective"]
    assert b["num_deductions"] < a["num_deductions"] / 2


def test_other_placements_fall_back_to_the_dense_sweeps(eng, orc):
    pb = tnf_gen.planted(500, 1500, 3)
    o = orc.fixpoint(pb)
    for kind in (abi.MEM_GLOBAL, abi.MEM_STORE_CLUSTER):
        with eng.Solver(pb, mem_kind=kind, fixpoint=abi.FP_WAC1_ACTIVE) as s:
            g = s.propagate()
        assert_same_store(g, o, kind)


# ---- snapshot ring for backtracking (copying instead of recomputation) --------------------------------------------------

@pytest.mark.parametrize("levels", ["0", "2", "3", "64"])
@pytest.mark.parametrize("fp", [abi.FP_WAC1, abi.FP_AC1_ACTIVE, abi.FP_WAC1_ACTIVE])
def test_snapshot_ring_sizes_keep_the_search_trace(eng, orc, monkeypatch, levels, fp):
    """0 = the reference's restore-from-root + replay; 2 and 3 levels force slot reuse and the fallback on almost every
    backtrack; 64 is the default. The oracle always recomputes: same nodes, failures, solutions, depth, optimum."""
    monkeypatch.setenv("TB_SNAPSHOT_LEVELS", levels)
    for seed in range(10):
        pb = tnf_gen.search_instance(seed) if seed < 7 else tnf_gen.random_net(18, 10, 4000 + seed, lo=-4, hi=4)
        for depth in (0, 3):
            o = orc.solve(pb, depth=depth)
            for kind in SHARED + [abi.MEM_GLOBAL]:
                with eng.Solver(pb, or_blocks=1, subproblems_power=depth, fixpoint=fp, mem_kind=kind) as s:
                    g = s.solve()
                for key in ("nodes", "fails", "solutions", "eps_solved_subproblems", "eps_skipped_subproblems", "depth_max"):
                    assert g["stats"][key] == o["stats"][key], (levels, seed, depth, kind, key, g["stats"][key], o["stats"][key])
                assert g["objective"] == o["objective"]


@pytest.mark.parametrize("levels", ["0", "2", "64"])
def test_snapshot_ring_on_a_deep_search(eng, monkeypatch, levels):
    """trains15 dives hundreds of levels deep: one block, 600 nodes, dense and active-set walk the same tree whatever the ring."""
    monkeypatch.setenv("TB_SNAPSHOT_LEVELS", levels)
    pb, info = golden_io.load_simplified_problem("trains15")
    res = []
    for fp in (abi.FP_WAC1, abi.FP_WAC1_ACTIVE):
        with eng.Solver(pb, or_blocks=1, subproblems_power=4, cutnodes=600, fixpoint=fp) as s:
            r = s.solve()
        res.append((r["stats"]["nodes"], r["stats"]["fails"], r["stats"]["solutions"], r["stats"]["depth_max"], r["objective"]))
    assert res[0] == res[1]
    assert res[0] == test_snapshot_ring_on_a_deep_search.expected.setdefault("trace", res[0])


test_snapshot_ring_on_a_deep_search.expected = {}


def test_small_networks_run_the_plain_sweeps_by_default(eng, orc, monkeypatch):
    """Under 128 chunks the *_ACTIVE kinds run the plain sweeps: no flag area is reserved next to the store."""
    pb = tnf_gen.planted(500, 1500, 3)              # about 50 chunks
    o = orc.fixpoint(pb)
    shared = {}
    for min_chunks in ("0", None):
        if min_chunks is None:
            monkeypatch.delenv("TB_ACTIVE_MIN_CHUNKS", raising=False)
        else:
            monkeypatch.setenv("TB_ACTIVE_MIN_CHUNKS", min_chunks)
        for fp in (abi.FP_WAC1, abi.FP_WAC1_ACTIVE):
            with eng.Solver(pb, fixpoint=fp, mem_kind=abi.MEM_TCN_SHARED) as s:
                shared[(min_chunks, fp)] = s.config()["shared_bytes"]
                g = s.propagate()
            assert_same_store(g, o, (min_chunks, fp))
    assert shared[("0", abi.FP_WAC1_ACTIVE)] > shared[("0", abi.FP_WAC1)]
    assert shared[(None, abi.FP_WAC1_ACTIVE)] == shared[(None, abi.FP_WAC1)] == shared[("0", abi.FP_WAC1)]
