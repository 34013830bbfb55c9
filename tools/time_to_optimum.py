#!/usr/bin/env python
"""Time-to-optimum / time-to-proof of the GPU dive-and-solve on a golden TNF fixture (BASELINE.json:
"time-to-optimum (s)"; the reference's `best_obj_time` and `solveTime`, include/statistics.hpp:346-369).

  python tools/time_to_optimum.py trains15 [--timeout-ms 20000] [--fp wac1]

Prints one JSON line: best objective, the known optimum of benchmarks/test_list.csv when the fixture has
one, seconds to the last improvement, seconds to the end of the search, whether the search was exhaustive.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import golden_io  # noqa: E402
from turbo_b200 import abi, engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--timeout-ms", type=int, default=20000)
    ap.add_argument("--fp", default="wac1")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--sub", type=int, default=-1, help="EPS depth (-1 = auto)")
    a = ap.parse_args()
    pb, info = (golden_io.load_simplified_problem(a.workload.split(':', 1)[1]) if a.workload.startswith('simplified:')
                else golden_io.load(a.workload))
    with engine.Solver(pb, device=a.device, timeout_ms=a.timeout_ms, subproblems_power=a.sub,
                       fixpoint=abi.FP_KINDS[a.fp]) as s:
        cfg = s.config()
        r = s.solve()
    st = r["stats"]
    secs = st["kernel_ms"] / 1e3
    out = {"workload": a.workload, "timeout_ms": a.timeout_ms, "fixpoint": a.fp,
           "memory_configuration": abi.MEM_NAMES.get(cfg["mem_kind"]), "num_blocks": cfg["num_blocks"],
           "subproblems_power": cfg["subproblems_power"],
           "has_solution": r["has_solution"], "exhaustive": r["exhaustive"],
           "best_objective": golden_io.user_objective(info, r["lb"], r["ub"]) if r["has_solution"] and info["objective_kind"] >= 0 else None,
           "known_optimum": info["expected"],
           "time_to_best_s": st["timers_ns"][abi.TIMER_LATEST_BEST_OBJ_FOUND] / 1e9,
           "solve_s": secs, "nodes": st["nodes"], "nodes_per_sec": st["nodes"] / secs if secs else None,
           "propagations_per_sec": st["num_deductions"] / secs if secs else None,
           "subproblems_solved": st.get("eps_solved_subproblems"), "subproblems_skipped": st.get("eps_skipped_subproblems")}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
