#!/usr/bin/env python
"""Benchmark sweep through the `turbo` driver (the author's workflow of README.md:41-50 — run a list of
instances through the solver entry MiniZinc would call and tabulate the %%%mzn-stat blocks).

  python tools/sweep.py [--timeout-ms 5000] [--gpus 1] [--fp wac1] [--out table.md] [instances ...]

An instance is a .fzn / .tnf path or the name of a golden fixture (tests/golden/<name>.npz, written to a temporary .tnf:
the FlatZinc sources of the reference are not on the GPU box). Without arguments: every golden fixture. Where MiniZinc is
installed the same sweep runs on .mzn/.dzn models with `minizinc --solver turbo_b200/minizinc/turbo.b200.release.msc -s`.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import golden_io  # noqa: E402

EXE = os.path.join(ROOT, "turbo_b200", "bin", "turbo")


def run_one(path, args):
    cmd = [EXE, "-s", "-t", str(args.timeout_ms), "-gpus", str(args.gpus), "-fp", args.fp] + (["-disable_simplify"] if args.disable_simplify else []) + [path]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=args.timeout_ms / 1000 + 120)
    st = dict(re.findall(r"^%%%mzn-stat: (\w+)=(.*)$", r.stdout, flags=re.M))
    out = r.stdout
    if "=====UNSATISFIABLE=====" in out:
        status = "unsat"
    elif "==========" in out:
        status = "optimal" if "objective" in st else "all solutions"
    elif "----------" in out:
        status = "feasible"
    else:
        status = "unknown"
    f = lambda k, d=None: (float(st[k]) if k in st else d)
    nodes, solve = f("nodes", 0.0), f("solveTime", 0.0)
    return {"rc": r.returncode, "status": status, "objective": int(st["objective"]) if "objective" in st else None,
            "solveTime": solve, "best_obj_time": f("best_obj_time"), "nodes": int(nodes), "initTime": f("initTime", 0.0),
            "nodes_per_sec": nodes / (solve - f("initTime", 0.0)) if solve - f("initTime", 0.0) > 0 else None,   # solving time without parsing / preprocessing
            "variables": int(f("variables", 0)), "propagators": int(f("propagators", 0)), "tcn_variables": int(f("tcn_variables", 0)),
            "tcn_constraints": int(f("tcn_constraints", 0)), "memory_configuration": st.get("memory_configuration", "").strip('"'),
            "num_blocks": int(f("num_blocks", 0)), "stderr": r.stderr.strip()[-200:]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("instances", nargs="*")
    ap.add_argument("--timeout-ms", type=int, default=5000)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--fp", default="wac1")
    ap.add_argument("--disable-simplify", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    names = args.instances or golden_io.names()
    rows = []
    with tempfile.TemporaryDirectory() as tmp:
        for name in names:
            expected = None
            if os.path.exists(name):
                path = name
            else:
                pb, info = golden_io.load(name)
                expected = info["expected"]
                path = os.path.join(tmp, name + ".tnf")
                golden_io.write_tnf(path, pb, info)
            r = run_one(path, args)
            r.update(instance=os.path.basename(name), expected=expected,
                     ok=(r["rc"] == 0 and (expected is None or (r["status"] == "optimal" and r["objective"] == expected))))
            rows.append(r)
            print(json.dumps(r), flush=True)
    lines = ["| instance | V (TNF → solved) | P (TNF → solved) | placement | status | objective | expected | solveTime s | time to best s | nodes | nodes/s |",
             "|---|---|---|---|---|---|---|---|---|---|---|"]
    for r in rows:
        lines.append("| %s | %d → %d | %d → %d | %s ×%d | %s | %s | %s | %.3f | %s | %d | %s |" % (
            r["instance"], r["tcn_variables"], r["variables"], r["tcn_constraints"], r["propagators"], r["memory_configuration"], r["num_blocks"],
            r["status"] + ("" if r["ok"] else " (!)"), r["objective"], r["expected"], r["solveTime"],
            "%.3f" % r["best_obj_time"] if r["best_obj_time"] is not None else "-", r["nodes"],
            "%.0f" % r["nodes_per_sec"] if r["nodes_per_sec"] else "-"))
    table = "\n".join(lines)
    if args.out:
        with open(args.out, "w") as f:
            f.write("`tools/sweep.py --timeout-ms %d --gpus %d --fp %s%s`\n\n" % (args.timeout_ms, args.gpus, args.fp, " --disable-simplify" if args.disable_simplify else ""))
            f.write(table + "\n")
    print(table)
    bad = [r["instance"] for r in rows if not r["ok"]]
    if bad:
        print("NOT OK:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
