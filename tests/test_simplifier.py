"""TNF simplifier (tb_model_simplify, turbo_b200/csrc/host/tnf_simplify.cpp; the reference's preprocess_tcn loop,
include/common_solving.hpp:538-565).  The simplifier is symbolic; its root fixpoint is a callback.  CPU tests drive it
with the oracle's fixpoint, GPU tests with the engine's (tb_fixpoint_on_device) and must arrive at the same network.

What is checked: the reduced network has the same optimum as the full one (the reference's known answers, and
brute force on random FlatZinc models), every solution of the reduced network expands to a point that satisfies every
ORIGINAL propagator and FlatZinc constraint, and the committed fixtures tests/golden/simplified/*.npz are what a fresh
run produces.
"""
import numpy as np
import pytest

from tests import golden_io
from tests.golden.make_simplified import oracle_fixpoint, simplified_model
from tests.test_frontend import gen_model
from turbo_b200 import abi
from turbo_b200.model import Model

ALL = golden_io.names()
WITH_ANSWER = [n for n in ALL if golden_io.load(n)[1]["expected"] is not None]
SLOW_ON_CPU = {"triangular9", "pat12", "pat13"}


def same_network(m, fx):
    a = golden_io.problem_arrays(m.problem)
    for k in ("lb", "ub", "props", "strat_meta", "strat_vars"):
        assert np.array_equal(a[k], fx[k]), k
    assert int(a["obj_var"]) == int(fx["obj_var"]) and m.user_objective_var == int(fx["user_obj_var"])


@pytest.mark.parametrize("name", ALL)
def test_fixture_is_what_the_simplifier_produces(name):
    same_network(simplified_model(name), golden_io.load_simplified(name))


@pytest.mark.parametrize("name", ALL)
def test_simplifier_shrinks_and_is_idempotent(name):
    m = simplified_model(name)
    st = m.simplify_stats
    assert st["vars_after"] <= st["vars_before"] and st["props_after"] <= st["props_before"]
    assert m.problem.nvars == st["vars_after"] and m.problem.nprops == st["props_after"]
    # the reduced network is at its root fixpoint and nothing in it is entailed or duplicated any more
    if m.problem.nprops:
        from oracle import oracle_py as orc
        r = orc.fixpoint(m.problem)
        assert not r["failed"] and np.array_equal(r["lb"], m.problem.lb) and np.array_equal(r["ub"], m.problem.ub)
        p = m.problem.props
        keys = set()
        for op, y, z in zip(p["op"], p["y"], p["z"]):
            if op in (abi.OP_ADD, abi.OP_MUL, abi.OP_MIN, abi.OP_MAX, abi.OP_EQ) and z < y:
                y, z = z, y
            keys.add((int(op), int(y), int(z)))
        assert len(keys) == m.problem.nprops
    again = m.simplify(oracle_fixpoint)          # a second call is a no-op
    assert again == st


@pytest.mark.parametrize("name", [n for n in WITH_ANSWER if n not in SLOW_ON_CPU])
def test_reduced_network_has_the_reference_optimum(name):
    from oracle import oracle_py as orc
    pb, info = golden_io.load(name)
    m = simplified_model(name)
    r = orc.solve(m.problem, depth=4, timeout_ms=60000)
    assert r["has_solution"] and r["exhaustive"]
    assert m.user_objective(r["lb"], r["ub"]) == info["expected"]
    # the solution expands to a point of the FULL network that satisfies every original propagator
    flb, fub = m.expand(r["lb"], r["ub"])
    assert len(flb) == pb.nvars
    assert np.all(flb >= pb.lb) and np.all(flb <= pb.ub)
    assert m.check_tnf(r["lb"]) == 0
    assert golden_io.user_objective(info, flb, fub) == info["expected"]


@pytest.mark.parametrize("seed", range(150))
def test_random_models_match_bruteforce_after_simplification(seed):
    from oracle import oracle_py as orc
    rng = np.random.default_rng(seed)
    text, best = gen_model(rng)
    m = Model.from_fzn_text(text)
    if m.root_failed:
        assert best is None, text
        return
    m.simplify(oracle_fixpoint)
    if m.root_failed:
        assert best is None, text
        return
    r = orc.solve(m.problem, depth=2)
    assert r["exhaustive"]
    assert r["has_solution"] == (best is not None), text
    if best is not None:
        assert m.user_objective(r["lb"], r["ub"]) == best, text
        assert m.check_solution(r["lb"]) == 0, text          # FlatZinc-level check on the expanded point
        assert m.check_tnf(r["lb"]) == 0, text


def test_equivalences_constants_and_useless_variables():
    # x = y (through int_eq), a duplicated sum, a constraint entailed at the root, a variable nobody constrains
    m = Model.from_fzn_text(
        "var 0..9: x :: output_var;\nvar 0..9: y :: output_var;\nvar 0..9: z :: output_var;\nvar 0..9: w :: output_var;\n"
        "var 0..20: s :: output_var;\nvar 0..20: t :: output_var;\n"
        "constraint int_eq(x, y);\nconstraint int_lin_eq([1,1,-1],[y,z,s],0);\nconstraint int_lin_eq([1,1,-1],[x,z,t],0);\n"
        "constraint int_le(z, 9);\nconstraint int_lin_le([-1],[s],-3);\nsolve minimize t;\n")
    full = m.problem.nvars
    st = m.simplify(oracle_fixpoint)
    assert st["vars_after"] < full and st["merged_variables"] >= 2 and st["eliminated_icse"] >= 1
    from oracle import oracle_py as orc
    r = orc.solve(m.problem, depth=0)
    assert r["has_solution"] and m.user_objective(r["lb"], r["ub"]) == 3
    assert m.check_solution(r["lb"]) == 0
    text = m.format_solution(r["lb"])
    vals = dict(line.rstrip(";").split(" = ") for line in text.strip().splitlines())
    assert vals["x"] == vals["y"] and vals["s"] == vals["t"] == "3" and vals["w"] == "0"


def test_functionally_defined_variables_are_computed_at_expansion():
    # d = x + z and e = (d <= 7) are only defined, never used: both propagators go, the values come back on expansion
    m = Model.from_fzn_text(
        "var 0..9: x :: output_var;\nvar 0..9: z :: output_var;\nvar 0..18: d :: output_var;\nvar bool: e :: output_var;\n"
        "constraint int_lin_eq([1,1,-1],[x,z,d],0);\nconstraint int_le_reif(d, 7, e);\nconstraint int_lin_le([-1,-2],[x,z],-11);\n"
        "solve minimize x;\n")
    st = m.simplify(oracle_fixpoint)
    assert st["eliminated_functional"] == 2
    from oracle import oracle_py as orc
    r = orc.solve(m.problem, depth=0)
    assert r["has_solution"] and m.check_solution(r["lb"]) == 0 and m.check_tnf(r["lb"]) == 0
    vals = dict(line.rstrip(";").split(" = ") for line in m.format_solution(r["lb"]).strip().splitlines())
    assert int(vals["d"]) == int(vals["x"]) + int(vals["z"]) and vals["e"] == ("true" if int(vals["d"]) <= 7 else "false")
    # a defined variable whose domain is tighter than the range of its definition is a constraint: it stays
    m = Model.from_fzn_text("var 0..9: x;\nvar 0..9: z;\nvar 0..5: d;\nconstraint int_lin_eq([1,1,-1],[x,z,d],0);\nsolve maximize x;\n")
    st = m.simplify(oracle_fixpoint)
    assert st["eliminated_functional"] == 0 and m.problem.nprops >= 1


def test_root_failure_is_detected_by_the_simplifier():
    m = Model.from_fzn_text("var 0..5: x;\nvar 0..5: y;\nconstraint int_lin_eq([1,1],[x,y],20);\nsolve satisfy;\n")
    m.simplify(oracle_fixpoint)
    assert m.root_failed


# ---- GPU: the engine's own fixpoint drives the simplifier --------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("name", ALL)
def test_gpu_driven_simplifier_matches_the_fixture(name):
    m = simplified_model(name, fixpoint="device")
    same_network(m, golden_io.load_simplified(name))


@pytest.mark.gpu
@pytest.mark.parametrize("name", WITH_ANSWER)
def test_gpu_solves_the_reduced_network_to_the_reference_optimum(name):
    from turbo_b200 import engine
    pb, info = golden_io.load(name)
    m = simplified_model(name, fixpoint="device")
    with engine.Solver(m.problem, device=0, timeout_ms=120000) as s:
        r = s.solve()
    assert r["has_solution"] and r["exhaustive"]
    assert m.user_objective(r["lb"], r["ub"]) == info["expected"]
    assert m.check_tnf(r["lb"]) == 0
    flb, fub = m.expand(r["lb"], r["ub"])
    assert golden_io.user_objective(info, flb, fub) == info["expected"]


def test_satisfaction_problems_and_strategies_after_simplification():
    from oracle import oracle_py as orc
    m = Model.from_fzn_text(
        "var 1..4: a :: output_var;\nvar 1..4: b :: output_var;\nvar 1..4: c :: output_var;\nvar 1..4: d :: output_var;\nvar bool: p;\n"
        "constraint int_ne(a, b);\nconstraint int_ne(b, c);\nconstraint int_ne(a, c);\nconstraint int_eq(c, d);\n"
        "constraint int_lin_le([1,1],[a,b],4);\nconstraint int_le_reif(a, b, p);\nconstraint bool_eq(p, true);\n"
        "solve :: int_search([d, c, b, a], input_order, indomain_max, complete) satisfy;\n")
    full_strategies = [list(vs) for _, _, vs in m.problem.strategies]
    st = m.simplify(oracle_fixpoint)
    assert st["merged_variables"] >= 1                       # c == d
    # the user's strategy survives on representatives, without duplicates, in the same order; the default stays last
    (vo, va, vs), (dvo, dva, dvs) = m.problem.strategies[0], m.problem.strategies[-1]
    assert va == abi.VAL_MAX and len(vs) == len(set(vs.tolist())) and 1 <= len(vs) <= len(full_strategies[0]) and len(dvs) == 0
    r = orc.solve(m.problem, depth=0)
    assert r["has_solution"] and m.check_solution(r["lb"]) == 0
    vals = dict(line.rstrip(";").split(" = ") for line in m.format_solution(r["lb"]).strip().splitlines())
    a, b, c, d = (int(vals[k]) for k in "abcd")
    assert len({a, b, c}) == 3 and c == d and a + b <= 4 and a <= b


@pytest.mark.parametrize("seed", range(120))
def test_random_tnf_networks_keep_status_and_optimum(seed, tmp_path):
    """Every operator (MUL, TDIV, TMOD, MIN, MAX, EQ, LEQ, ADD) on small random networks, straight at the TNF level: the
    reduced network has the same satisfiability and the same optimum as the full one, and its solutions expand to points
    that satisfy every original propagator."""
    from oracle import oracle_py as orc
    from tests import tnf_gen
    pb = tnf_gen.search_instance(seed // 3 % 12) if seed % 3 == 0 else tnf_gen.random_net(14, 9, 7000 + seed, lo=-4, hi=4)
    o = orc.solve(pb, depth=0, timeout_ms=60000)
    info = dict(objective_kind=0 if pb.obj_var >= 0 else -1, user_obj_var=pb.obj_var)
    path = str(tmp_path / "net.tnf")
    golden_io.write_tnf(path, pb, info)
    m = Model.from_tnf(path)
    m.simplify(oracle_fixpoint)
    if m.root_failed:
        assert not o["has_solution"]
        return
    r = orc.solve(m.problem, depth=0, timeout_ms=60000)
    assert r["exhaustive"] and o["exhaustive"]
    assert r["has_solution"] == o["has_solution"], seed
    if o["has_solution"]:
        if pb.obj_var >= 0:
            assert m.user_objective(r["lb"], r["ub"]) == o["objective"], seed
        assert m.check_tnf(r["lb"]) == 0, seed
