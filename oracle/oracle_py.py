"""ctypes binding of oracle/libtnf_oracle.so — TEST INFRASTRUCTURE ONLY.

May be imported by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
`--impl reference` legs, always as the checker or the reported CPU baseline, never as the product
path.  See oracle/tnf_oracle.h for the "parity unpinned" statement.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from turbo_b200 import abi  # noqa: E402

_LIB = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libtnf_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        i32p = C.POINTER(C.c_int32)
        L.tbo_deduce.argtypes = [C.POINTER(abi.TbProp), i32p, i32p, i32p]
        L.tbo_deduce.restype = C.c_int
        L.tbo_ask.argtypes = [C.POINTER(abi.TbProp), i32p, i32p]
        L.tbo_ask.restype = C.c_int
        L.tbo_fixpoint.argtypes = [C.POINTER(abi.TbProblem), i32p, i32p, i32p, i32p, C.POINTER(C.c_uint64)]
        L.tbo_fixpoint.restype = C.c_int64
        L.tbo_dive.argtypes = [C.POINTER(abi.TbProblem), C.c_uint64, C.c_int32, i32p, i32p, i32p, i32p]
        L.tbo_dive.restype = C.c_int
        L.tbo_solve.argtypes = [C.POINTER(abi.TbProblem), C.c_int32, C.c_uint64, C.c_uint64, C.c_int32,
                                i32p, i32p, i32p, i32p, i32p, C.POINTER(abi.TbStats)]
        L.tbo_solve.restype = C.c_int
        L.tbo_solve_shard.argtypes = [C.POINTER(abi.TbProblem), C.c_int32, C.c_uint64, C.c_uint64, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_int32,
                                      i32p, i32p, i32p, i32p, i32p, C.POINTER(abi.TbStats)]
        L.tbo_solve_shard.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def deduce(op, x, y, z, lb, ub):
    """One narrowing step on copies of (lb, ub). Returns (changed, failed, lb, ub)."""
    lb = np.array(lb, dtype=np.int32)
    ub = np.array(ub, dtype=np.int32)
    p = abi.TbProp(op, x, y, z)
    failed = C.c_int32(0)
    ch = lib().tbo_deduce(C.byref(p), _p(lb), _p(ub), C.byref(failed))
    return bool(ch), bool(failed.value), lb, ub


def ask(op, x, y, z, lb, ub):
    lb = np.ascontiguousarray(lb, dtype=np.int32)
    ub = np.ascontiguousarray(ub, dtype=np.int32)
    p = abi.TbProp(op, x, y, z)
    return bool(lib().tbo_ask(C.byref(p), _p(lb), _p(ub)))


def fixpoint(problem, lb=None, ub=None, order=None):
    """Gauss-Seidel fixpoint. Returns dict(lb, ub, failed, sweeps, num_deductions)."""
    lb = np.array(problem.lb if lb is None else lb, dtype=np.int32)
    ub = np.array(problem.ub if ub is None else ub, dtype=np.int32)
    failed = C.c_int32(0)
    nd = C.c_uint64(0)
    o = None
    if order is not None:
        order = np.ascontiguousarray(order, dtype=np.int32)
        o = _p(order)
    sweeps = lib().tbo_fixpoint(C.byref(problem.c), _p(lb), _p(ub), C.byref(failed), o, C.byref(nd))
    return dict(lb=lb, ub=ub, failed=bool(failed.value), sweeps=int(sweeps), num_deductions=int(nd.value))


def dive(problem, idx, depth):
    lb = np.zeros(problem.nvars, np.int32)
    ub = np.zeros(problem.nvars, np.int32)
    rem = C.c_int32(0)
    kind = C.c_int32(0)
    rc = lib().tbo_dive(C.byref(problem.c), idx, depth, _p(lb), _p(ub), C.byref(rem), C.byref(kind))
    assert rc == 0, rc
    return dict(lb=lb, ub=ub, remaining_depth=rem.value, leaf_kind=kind.value)


def solve(problem, depth=0, cutnodes=0, timeout_ms=0, nthreads=1, rank=0, world=1, initial_bound=abi.POS_INF):
    n = max(1, problem.nvars)
    lb = np.zeros(n, np.int32)
    ub = np.zeros(n, np.int32)
    has = C.c_int32(0)
    exh = C.c_int32(0)
    st = abi.TbStats()
    rc = lib().tbo_solve_shard(C.byref(problem.c), depth, cutnodes, timeout_ms, nthreads, rank, world, initial_bound,
                               None, _p(lb), _p(ub), C.byref(has), C.byref(exh), C.byref(st))
    assert rc == 0, rc
    obj = None
    if has.value and problem.obj_var >= 0:
        obj = int(lb[problem.obj_var])
    return dict(lb=lb[:problem.nvars], ub=ub[:problem.nvars], has_solution=bool(has.value),
                exhaustive=bool(exh.value), objective=obj, stats=st.as_dict())
