#!/usr/bin/env python
"""Independent cross-check of an optimum with z3 ("z3-solver for optimum/unsat on small and
medium instances"; BASELINE.md: the optima of the three headline instances "must be established by our own
exhaustive runs and cross-checked").

The TNF network of a golden fixture is restated as z3 integer constraints, one per propagator, with z3's own
semantics of + * min max = <= (nothing of the engine or of the oracle is involved), and two questions are asked:
  * is there a point with objective <= claimed optimum?      (must be sat, and the model must evaluate to it)
  * is there a point with objective <= claimed optimum - 1?  (must be unsat)

  python tools/z3_check.py accap_a3 --optimum 38 [--simplified] [--timeout 3600]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import golden_io  # noqa: E402
from turbo_b200 import abi  # noqa: E402


def build(pb, z3):
    xs = [z3.Int("v%d" % i) for i in range(pb.nvars)]
    cons = []
    for i in range(pb.nvars):
        lo, hi = int(pb.lb[i]), int(pb.ub[i])
        if lo != abi.NEG_INF:
            cons.append(xs[i] >= lo)
        if hi != abi.POS_INF:
            cons.append(xs[i] <= hi)

    def tdiv(a, b):      # truncated division from z3's floor/Euclidean one
        q = a / b
        return z3.If(z3.And(a % b != 0, a < 0), z3.If(b > 0, q + 1, q - 1), q)

    for p in pb.props:
        op, x, y, z = int(p["op"]), xs[int(p["x"])], xs[int(p["y"])], xs[int(p["z"])]
        if op == abi.OP_ADD:
            cons.append(x == y + z)
        elif op == abi.OP_MUL:
            cons.append(x == y * z)
        elif op == abi.OP_TDIV:
            cons.append(z != 0)
            cons.append(x == tdiv(y, z))
        elif op == abi.OP_TMOD:
            cons.append(z != 0)
            cons.append(x == y - z * tdiv(y, z))
        elif op == abi.OP_MIN:
            cons.append(x == z3.If(y <= z, y, z))
        elif op == abi.OP_MAX:
            cons.append(x == z3.If(y >= z, y, z))
        elif op == abi.OP_EQ:
            cons.append((x == 1) == (y == z))
        elif op == abi.OP_LEQ:
            cons.append((x == 1) == (y <= z))
        else:
            raise ValueError(op)
    return xs, cons


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name")
    ap.add_argument("--optimum", type=int, required=True, help="claimed optimum, in the USER's objective (as printed by objective=)")
    ap.add_argument("--simplified", action="store_true")
    ap.add_argument("--timeout", type=int, default=3600)
    args = ap.parse_args()
    import z3
    pb, info = (golden_io.load_simplified_problem if args.simplified else golden_io.load)(args.name)
    # the TNF always minimises obj_var; a maximised user objective is its negation (common_solving.hpp:489-510)
    tnf_opt = args.optimum if info["objective_kind"] == 0 else -args.optimum
    xs, cons = build(pb, z3)
    out = {"instance": args.name, "simplified": args.simplified, "claimed_optimum": args.optimum, "nvars": pb.nvars, "nprops": pb.nprops,
           "z3": z3.get_version_string()}
    for what, bound, want in (("feasible_at_optimum", tnf_opt, "sat"), ("nothing_better", tnf_opt - 1, "unsat")):
        s = z3.Solver()
        s.set("timeout", args.timeout * 1000)
        s.add(cons)
        s.add(xs[pb.obj_var] <= bound)
        t = time.time()
        r = str(s.check())
        out[what] = {"result": r, "expected": want, "seconds": round(time.time() - t, 1)}
        if r == "sat":
            out[what]["objective_of_model"] = s.model().eval(xs[pb.obj_var]).as_long()
        print(json.dumps(out), flush=True)
    out["confirmed"] = out["feasible_at_optimum"]["result"] == "sat" and out["nothing_better"]["result"] == "unsat"
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
