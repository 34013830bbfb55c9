"""Regenerates tests/golden/*.npz from the reference's FlatZinc inputs.

Run here (the container that has /root/reference); the GPU box only ever sees the committed
fixtures. For every instance: the TNF produced by our C++ front-end (lb, ub, props, strategies,
objective), the reference's expected optimum (benchmarks/test_list.csv) where one exists, and the
sha256 of the oracle's root fixpoint (a regression pin for front-end + oracle).
"""
import csv
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle_py as orc  # noqa: E402
from turbo_b200.model import Model  # noqa: E402

REF = "/root/reference/benchmarks"


def root_sha(pb):
    r = orc.fixpoint(pb)
    h = hashlib.sha256()
    h.update(b"F" if r["failed"] else b"-")
    if not r["failed"]:
        h.update(r["lb"].tobytes())
        h.update(r["ub"].tobytes())
    return h.hexdigest(), r["failed"]


def dump(name, path, expected):
    m = Model.from_fzn(path)
    pb = m.problem
    sha, failed = root_sha(pb)
    props = np.stack([pb.props[f] for f in ("op", "x", "y", "z")], axis=1).astype(np.int32) if pb.nprops else np.zeros((0, 4), np.int32)
    strat_meta = np.array([(vo, va, len(vs)) for vo, va, vs in pb.strategies], dtype=np.int32).reshape(-1, 3)
    strat_vars = np.concatenate([vs for _, _, vs in pb.strategies] + [np.zeros(0, np.int32)]).astype(np.int32)
    meta = np.array([pb.obj_var, pb.c.has_eps_strategy, m.objective_kind, m.user_objective_var,
                     0 if expected is None else expected, 0 if expected is None else 1, int(failed)], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), lb=pb.lb, ub=pb.ub, props=props, strat_meta=strat_meta,
                        strat_vars=strat_vars, meta=meta, root_sha=np.array(sha))
    print(f"{name:28s} V={pb.nvars:6d} P={pb.nprops:6d} expected={expected} root={sha[:12]}")


def main():
    for path, exp in csv.reader(open(os.path.join(REF, "test_list.csv"))):
        if path.endswith(".xml"):
            continue            # XCSP3: out of scope
        name = os.path.basename(path)[:-4]
        dump(name, os.path.join("/root/reference", path), int(exp))
    for name in ("accap_a3", "trains15", "example_wordpress7_500"):
        dump(name, os.path.join(REF, name + ".fzn"), None)
    for name in ("bigdom", "valve6"):            # unsolved_bugs_data: no expected answer; valve6 has set variables
        dump(name, os.path.join(REF, "unsolved_bugs_data", name + ".fzn"), None)


if __name__ == "__main__":
    main()
