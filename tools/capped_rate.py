#!/usr/bin/env python
"""Sustained node rate of a real search in a fixed regime: accap_a3 (simplified) with the objective capped at 104 (below the
best known value: no incumbent ever moves), one GPU, fixed budget.   python tools/capped_rate.py [budget_ms]
Run with TB_SHAPE_V1=1 for round 1's placement and block shape."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import golden_io                                    # noqa: E402
from turbo_b200 import abi, engine as eng                      # noqa: E402

budget = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
pb, info = golden_io.load_simplified_problem("accap_a3")
ub = np.array(pb.ub, np.int32).copy()
ub[pb.obj_var] = min(int(ub[pb.obj_var]), 104)
capped = abi.Problem(np.array(pb.lb, np.int32), ub, np.stack([pb.props[k] for k in ("op", "x", "y", "z")], axis=1).astype(np.int32),
                     pb.strategies, obj_var=pb.obj_var, has_eps_strategy=int(pb.c.has_eps_strategy))
for power in (12, -1):
    with eng.Solver(capped, subproblems_power=power, timeout_ms=budget) as s:
        r = s.solve()
        st = r["stats"]
        print(json.dumps(dict(workload="simplified:accap_a3, objective <= 104", shape_v1=os.environ.get("TB_SHAPE_V1", "0"), budget_ms=budget,
                              subproblems_power=st["subproblems_power"], blocks=st["num_blocks"], threads_per_block=st["threads_per_block"], mem_kind=st["mem_kind"],
                              nodes=st["nodes"], nodes_per_sec=round(st["nodes"] / (st["kernel_ms"] * 1e-3)), propagations_per_sec=round(st["num_deductions"] / (st["kernel_ms"] * 1e-3)),
                              has_solution=r["has_solution"], exhaustive=r["exhaustive"], split=st["eps_split_subproblems"], parts=st["eps_split_parts_solved"])), flush=True)
