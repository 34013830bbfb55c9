"""The host layout pass (turbo_b200/csrc/layout.cpp) on its own: propagator classes, constant folding,
bank-aware variable placement. No GPU needed: tb_layout_describe is pure host code."""
import numpy as np
import pytest

from tests import golden_io, tnf_gen
from turbo_b200 import abi, engine

NI, PI = abi.NEG_INF, abi.POS_INF


@pytest.mark.parametrize("name", ["trains15", "accap_a3", "example_wordpress7_500", "pat1", "sudoku_opt3"])
def test_placement_is_a_permutation_and_lowers_bank_conflicts(name):
    pb, _ = golden_io.load(name)
    ident = engine.layout_describe(pb, nbanks=0)
    banked = engine.layout_describe(pb, nbanks=16)
    assert ident["identity"] and np.array_equal(ident["slot_of"], np.arange(pb.nvars))
    s = banked["slot_of"]
    assert len(np.unique(s)) == pb.nvars and s.min() >= 0 and s.max() < banked["nslots"]
    assert banked["nslots"] % 16 == 0 and banked["nslots"] - pb.nvars < 32
    # every propagator lands in exactly one class, whatever the placement
    assert sum(ident["classes"].values()) == pb.nprops == sum(banked["classes"].values())
    assert ident["classes"] == banked["classes"]
    # the bank model: 1.0 = conflict free; the placement must not be worse than the ternariser's numbering
    assert 1.0 <= banked["wavefronts_per_load"] <= ident["wavefronts_per_load"] + 1e-9
    if pb.nprops > 1000:
        assert banked["wavefronts_per_load"] < 1.25
    # folded constants cost no load: fewer than 3 interval loads per propagator on these models
    assert banked["loads_per_sweep"] <= 3 * pb.nprops


def test_classes_follow_the_root_domains():
    #        0       1       2       3          4          5           6            7
    lb = [0,      1,      5,      0,         -10,       NI,         0,           3_000_000]
    ub = [0,      1,      5,      1,         10,        PI,         2 ** 30,     3_000_000]
    P = lambda op, x, y, z: (op, x, y, z)
    cases = [
        (P(abi.OP_ADD, 4, 4, 4), "add_s"), (P(abi.OP_ADD, 2, 4, 4), "add_xk"), (P(abi.OP_ADD, 4, 4, 2), "add_zk"),
        (P(abi.OP_ADD, 4, 2, 4), "add_zk"),           # y constant: operands swapped
        (P(abi.OP_ADD, 4, 5, 4), "add_g"), (P(abi.OP_ADD, 4, 6, 4), "add_g"),      # infinite / huge bounds
        (P(abi.OP_ADD, 4, 4, 7), "add_s"),            # constant too large for a 21-bit field: stays a variable
        (P(abi.OP_LEQ, 1, 4, 4), "leq_t"), (P(abi.OP_LEQ, 0, 4, 4), "leq_f"), (P(abi.OP_LEQ, 3, 4, 2), "leq_zk"),
        (P(abi.OP_LEQ, 3, 4, 4), "leq_s"), (P(abi.OP_LEQ, 3, 5, 4), "leq_g"),
        (P(abi.OP_EQ, 1, 4, 4), "eq_t"), (P(abi.OP_EQ, 0, 4, 4), "eq_f"), (P(abi.OP_EQ, 3, 4, 2), "eq_zk"),
        (P(abi.OP_EQ, 3, 2, 4), "eq_zk"), (P(abi.OP_EQ, 3, 4, 4), "eq_s"), (P(abi.OP_EQ, 3, 4, 5), "eq_g"),
        (P(abi.OP_MUL, 4, 4, 4), "mul"), (P(abi.OP_TDIV, 4, 4, 4), "tdiv"), (P(abi.OP_TMOD, 4, 4, 4), "tmod"),
        (P(abi.OP_MIN, 4, 4, 4), "min"), (P(abi.OP_MAX, 4, 4, 4), "max"),
    ]
    for prop, want in cases:
        pb = abi.Problem(lb, ub, np.array([prop], np.int32))
        got = {k: v for k, v in engine.layout_describe(pb, 16)["classes"].items() if v}
        assert got == {want: 1}, (prop, got, want)


def test_random_networks_layout_invariants():
    for seed in range(10):
        pb = tnf_gen.planted(200 + 50 * seed, 600 + 100 * seed, seed)
        d = engine.layout_describe(pb, 16)
        assert sum(d["classes"].values()) == pb.nprops
        assert len(np.unique(d["slot_of"])) == pb.nvars
        assert d["nchunks"] >= (pb.nprops + 31) // 32 // 1 - 0 and d["nchunks"] <= pb.nprops // 32 + len(d["classes"])


def test_empty_and_tiny_problems():
    pb = abi.Problem([0, 1], [0, 1], np.zeros((0, 4), np.int32))
    d = engine.layout_describe(pb, 16)
    assert d["nchunks"] == 0 and sum(d["classes"].values()) == 0
    pb = abi.Problem([0, 0, 0], [5, 5, 5], np.array([(abi.OP_ADD, 0, 1, 2)], np.int32))
    d = engine.layout_describe(pb, 16)
    assert d["nchunks"] == 1 and d["classes"]["add_s"] == 1


@pytest.mark.parametrize("name", ["trains15", "accap_a3", "pat1", "sudoku_opt3"])
def test_watch_lists_cover_every_loaded_operand(name):
    """The active-set fixpoint re-evaluates exactly the chunks on a slot's watch list when the slot moves: the chunk of
    every propagator must be on the list of every NON-CONSTANT variable it mentions (operands that are constants at the
    root are folded into the device word and never move), lists are sorted and duplicate free."""
    from tests import golden_io
    from turbo_b200 import engine
    pb, _ = golden_io.load_simplified_problem(name)
    off, lst, slot_of, chunk_of = engine.layout_watch_lists(pb, 16)
    assert len(set(slot_of.tolist())) == pb.nvars and off[-1] == len(lst)
    watch = [set(lst[off[s]:off[s + 1]].tolist()) for s in range(len(off) - 1)]
    for s in range(len(off) - 1):
        seg = lst[off[s]:off[s + 1]]
        assert all(seg[i] < seg[i + 1] for i in range(len(seg) - 1))
    p = pb.props
    fixed = pb.lb == pb.ub
    missing = 0
    for i in range(pb.nprops):
        ch = int(chunk_of[i])
        assert ch >= 0
        for v in (int(p["x"][i]), int(p["y"][i]), int(p["z"][i])):
            if not fixed[v] and ch not in watch[int(slot_of[v])]:
                missing += 1
    assert missing == 0


def test_cluster_placement_keeps_most_operand_loads_in_the_evaluating_cta():
    """STORE_CLUSTER: variables that occur together share a CTA and a CTA evaluates the chunks whose operands mostly
    live in it; only the cut goes through DSMEM. Striping (slot = variable index) keeps 1 / C of the loads local."""
    import ctypes as C
    from tests import golden_io
    L = engine.lib()
    L.tb_layout_cluster_locality.argtypes = [C.POINTER(abi.TbProblem), C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.tb_layout_cluster_locality.restype = C.c_int
    pb, _ = golden_io.load("example_wordpress7_500")
    placed, striped = C.c_double(0), C.c_double(0)
    assert L.tb_layout_cluster_locality(C.byref(pb.c), 4, 32, C.byref(placed), C.byref(striped)) == 0
    assert 0.2 < striped.value < 0.3 and placed.value > 0.65
    rnd = tnf_gen.planted(4000, 12000, 3)            # a random network has no locality beyond "two operands of three"
    assert L.tb_layout_cluster_locality(C.byref(rnd.c), 4, 32, C.byref(placed), C.byref(striped)) == 0
    assert placed.value > striped.value + 0.1 and striped.value < 0.3
