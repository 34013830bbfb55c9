#!/usr/bin/env python
"""Does sharing the tail-splitting pool between linked GPUs pay?  Two measurements that do not depend on when the search
happens to find its best solution:
  A. accap_a3 (simplified) with the objective capped BELOW the best known value (<= 104): no incumbent ever moves, every
     run is in the same regime from t = 0, and nodes in a fixed budget compare directly;
  B. instances that terminate (known answers): wall time to exhaust on 1 and on N GPUs.
Each on N GPUs with TB_SHARE_SPLIT=1 and =0 (read at link time), and on one GPU.   python tools/tail_share.py [N]"""
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import golden_io                                    # noqa: E402
from turbo_b200 import abi, engine as eng                      # noqa: E402


def run_all(solvers):
    res = [None] * len(solvers)

    def run(g):
        res[g] = solvers[g].solve()
    th = [threading.Thread(target=run, args=(g,)) for g in range(len(solvers))]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    return res, (time.perf_counter() - t0) * 1e3


def measure(pb, n, share, reps, **kw):
    os.environ["TB_SHARE_SPLIT"] = share
    solvers = [eng.Solver(pb, device=g, gpu_rank=g, gpu_world=n, **kw) for g in range(n)]
    if n > 1:
        eng.link_peers(solvers)
    best = None
    for _ in range(reps):
        res, wall = run_all(solvers)
        rec = dict(wall_ms=round(wall, 1), kernel_ms=round(max(r["stats"]["kernel_ms"] for r in res), 1),
                   nodes=sum(r["stats"]["nodes"] for r in res), nodes_per_gpu=[r["stats"]["nodes"] for r in res],
                   exhaustive=all(r["exhaustive"] for r in res), has_solution=any(r["has_solution"] for r in res),
                   objective=min([r["objective"] for r in res if r["has_solution"]], default=None),
                   stolen=sum(r["stats"]["eps_stolen_subproblems"] for r in res), split=sum(r["stats"]["eps_split_subproblems"] for r in res),
                   parts=sum(r["stats"]["eps_split_parts_solved"] for r in res))
        if best is None or rec["kernel_ms"] < best["kernel_ms"]:
            best = rec
    for s in solvers:
        s.close()
    return best


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else min(eng.device_count(), 8)
    budget = int(os.environ.get("TAIL_BUDGET_MS", "8000"))
    pb, info = golden_io.load_simplified_problem("accap_a3")
    ub = np.array(pb.ub, np.int32).copy()
    ub[pb.obj_var] = min(int(ub[pb.obj_var]), 104)
    capped = abi.Problem(np.array(pb.lb, np.int32), ub, np.stack([pb.props[k] for k in ("op", "x", "y", "z")], axis=1).astype(np.int32),
                         pb.strategies, obj_var=pb.obj_var, has_eps_strategy=int(pb.c.has_eps_strategy))
    for g, share in ((1, "1"), (n, "1"), (n, "0")):
        r = measure(capped, g, share, 1, subproblems_power=12, timeout_ms=budget)
        r["nodes_per_sec"] = round(r["nodes"] / (r["kernel_ms"] * 1e-3))
        print(json.dumps(dict(workload="simplified:accap_a3, objective <= 104", gpus=g, share_split=share, budget_ms=budget, **r)), flush=True)
    if os.environ.get("TAIL_ONLY_CAPPED"):
        return
    for name in ("triangular9", "pat3", "pat10", "pat1"):
        pb, info = golden_io.load(name)
        for g, share in ((1, "1"), (n, "1"), (n, "0")):
            r = measure(pb, g, share, 3, subproblems_power=12, timeout_ms=60000)
            print(json.dumps(dict(workload=name, gpus=g, share_split=share, **r)), flush=True)


if __name__ == "__main__":
    main()
