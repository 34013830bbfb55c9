"""Regenerates tests/golden/simplified/*.npz: what the TNF simplifier (tb_model_simplify) leaves of every golden
network when its root fixpoint is the CPU oracle's.  The GPU tests re-run the simplifier with the engine's fixpoint
(tb_fixpoint_on_device) and must arrive at the same arrays.  Needs no FlatZinc source: starts from tests/golden/*.npz.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle_py as orc  # noqa: E402
from tests import golden_io  # noqa: E402
from turbo_b200.model import Model  # noqa: E402


def oracle_fixpoint(pb):
    r = orc.fixpoint(pb)
    return r["lb"], r["ub"], r["failed"]


def simplified_model(name, fixpoint=oracle_fixpoint, **kw):
    pb, info = golden_io.load(name)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, name + ".tnf")
        golden_io.write_tnf(path, pb, info)
        m = Model.from_tnf(path)
    m.simplify(fixpoint, **kw)
    return m


def main():
    os.makedirs(os.path.join(HERE, "simplified"), exist_ok=True)
    for name in golden_io.names():
        m = simplified_model(name)
        a = golden_io.problem_arrays(m.problem)
        st = m.simplify_stats
        full = m.num_full_variables
        ident = np.arange(m.problem.nvars, dtype=np.int32)
        flb, _ = m.expand(ident)              # expansion of the identity store = the variable map (eliminated: root lb)
        np.savez_compressed(os.path.join(HERE, "simplified", name + ".npz"), user_obj_var=np.array(m.user_objective_var, np.int32),
                            stats=np.array([st[k] for k in sorted(st)], np.int32), **a)
        print(f"{name:28s} V {st['vars_before']:6d} -> {st['vars_after']:6d}   P {st['props_before']:6d} -> {st['props_after']:6d}"
              f"   iterations {st['iterations']} full={full}")


if __name__ == "__main__":
    main()
