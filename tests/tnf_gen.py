"""Seeded random TNF networks for the parity tests (small relatives of BASELINE config 5,
SURVEY.md §8d: planted solution, mixed operators, constants 0/1/2 as variables 0/1/2)."""
import numpy as np

from turbo_b200 import abi

DEFAULT_MIX = {abi.OP_ADD: 0.40, abi.OP_LEQ: 0.25, abi.OP_EQ: 0.15, abi.OP_MIN: 0.05, abi.OP_MAX: 0.05,
               abi.OP_MUL: 0.10}


def _tdiv(a, b):
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b > 0) else -q


def _apply(op, a, b):
    if op == abi.OP_ADD:
        return a + b
    if op == abi.OP_MUL:
        return a * b
    if op == abi.OP_TDIV:
        return None if b == 0 else _tdiv(a, b)
    if op == abi.OP_TMOD:
        return None if b == 0 else a - b * _tdiv(a, b)
    if op == abi.OP_MIN:
        return min(a, b)
    if op == abi.OP_MAX:
        return max(a, b)
    if op == abi.OP_EQ:
        return int(a == b)
    if op == abi.OP_LEQ:
        return int(a <= b)
    raise ValueError(op)


def planted(nvars, nprops, seed, spread=50, slack=32, mix=None, singleton_frac=0.10, objective=False,
            strategies=None):
    """A satisfiable network: every propagator holds on a hidden assignment, domains contain it."""
    rng = np.random.default_rng(seed)
    mix = mix or DEFAULT_MIX
    ops = list(mix.keys())
    probs = np.array([mix[o] for o in ops], dtype=float)
    probs /= probs.sum()
    nvars = max(nvars, 4)
    s = rng.integers(-spread, spread + 1, size=nvars)
    nbool = max(1, nvars // 10)
    boolv = rng.choice(np.arange(3, nvars), size=min(nbool, nvars - 3), replace=False)
    s[boolv] = rng.integers(0, 2, size=len(boolv))
    s[0], s[1], s[2] = 0, 1, 2
    by_value = {}
    for v in range(nvars):
        by_value.setdefault(int(s[v]), []).append(v)
    props = []
    tries = 0
    while len(props) < nprops and tries < nprops * 200:
        tries += 1
        op = int(rng.choice(ops, p=probs))
        y, z = (int(t) for t in rng.integers(0, nvars, size=2))
        if op == abi.OP_MUL and (abs(s[y]) > 20 or abs(s[z]) > 20):
            continue
        r = _apply(op, int(s[y]), int(s[z]))
        if r is None or r not in by_value:
            continue
        cands = by_value[r]
        x = int(cands[rng.integers(0, len(cands))])
        props.append((op, x, y, z))
    lb = np.empty(nvars, np.int64)
    ub = np.empty(nvars, np.int64)
    for v in range(nvars):
        if rng.random() < singleton_frac:
            lb[v] = ub[v] = s[v]
        else:
            lb[v] = s[v] - rng.integers(0, slack + 1)
            ub[v] = s[v] + rng.integers(0, slack + 1)
    isbool = np.zeros(nvars, bool)
    isbool[boolv] = True
    lb[isbool] = np.maximum(lb[isbool], 0)
    ub[isbool] = np.minimum(ub[isbool], 1)
    for k in range(3):
        lb[k] = ub[k] = k
    # the result of a reified comparison is a 0..1 variable (precondition of TB_OP_EQ / TB_OP_LEQ)
    for op, x, _, _ in props:
        if op in (abi.OP_EQ, abi.OP_LEQ):
            lb[x] = max(lb[x], 0)
            ub[x] = min(ub[x], 1)
    obj = -1
    if objective:
        obj = int(rng.integers(3, nvars))
    pb = abi.Problem(lb, ub, np.array(props, dtype=np.int32).reshape(-1, 4), strategies, obj_var=obj)
    pb.planted = s
    return pb


def random_net(nvars, nprops, seed, lo=-6, hi=6, objective=True, mix=None, strategies=None):
    """No planted solution: small domains, arbitrary propagators; often unsatisfiable or with a
    non-trivial optimum. Used for search parity (status + optimum + node counts)."""
    rng = np.random.default_rng(seed)
    mix = mix or {abi.OP_ADD: 0.35, abi.OP_LEQ: 0.25, abi.OP_EQ: 0.2, abi.OP_MIN: 0.05, abi.OP_MAX: 0.05,
                  abi.OP_MUL: 0.06, abi.OP_TDIV: 0.02, abi.OP_TMOD: 0.02}
    ops = list(mix.keys())
    probs = np.array([mix[o] for o in ops], dtype=float)
    probs /= probs.sum()
    nvars = max(nvars, 6)
    lb = np.empty(nvars, np.int64)
    ub = np.empty(nvars, np.int64)
    nbool = max(2, nvars // 4)
    for v in range(nvars):
        if 3 <= v < 3 + nbool:
            lb[v], ub[v] = 0, 1
        else:
            a, b = sorted(rng.integers(lo, hi + 1, size=2))
            lb[v], ub[v] = a, b
    for k in range(3):
        lb[k] = ub[k] = k
    props = []
    for _ in range(nprops):
        op = int(rng.choice(ops, p=probs))
        y, z = (int(t) for t in rng.integers(0, nvars, size=2))
        if op in (abi.OP_EQ, abi.OP_LEQ):
            x = int(rng.choice([0, 1, 1] + list(range(3, 3 + nbool))))
        else:
            x = int(rng.integers(3 + nbool, nvars))
        props.append((op, x, y, z))
    obj = int(rng.integers(3 + nbool, nvars)) if objective else -1
    return abi.Problem(lb, ub, np.array(props, dtype=np.int32).reshape(-1, 4), strategies, obj_var=obj)


def search_instance(seed):
    """A satisfiable optimisation instance whose proof of optimality needs a real search tree
    (tens to hundreds of nodes, several improving solutions, backtracking)."""
    if seed % 2 == 0:
        return planted(60, 60, 6000 + seed, spread=12, slack=12, objective=True, singleton_frac=0.05)
    return planted(80, 70, 6000 + seed, spread=10, slack=16, objective=True, singleton_frac=0.05)
