"""ctypes mirror of include/turbo_b200.h (plain C ABI; no torch types cross it).

This module only declares the structs and a helper that owns the host buffers of a
`tb_problem`; it is shared by the product binding (`turbo_b200.engine`) and by the oracle
binding under `oracle/` (test infrastructure).
"""
import ctypes as C

import numpy as np

NEG_INF = -(2 ** 31)
POS_INF = 2 ** 31 - 1

OP_ADD, OP_MUL, OP_TDIV, OP_TMOD, OP_MIN, OP_MAX, OP_EQ, OP_LEQ = range(8)
OP_NAMES = ["ADD", "MUL", "TDIV", "TMOD", "MIN", "MAX", "EQ", "LEQ"]
VAR_INPUT_ORDER, VAR_FIRST_FAIL, VAR_ANTI_FIRST_FAIL, VAR_SMALLEST, VAR_LARGEST = range(5)
VAL_MIN, VAL_MAX, VAL_SPLIT, VAL_REVERSE_SPLIT = range(4)
FP_AC1, FP_WAC1, FP_AC1_ACTIVE, FP_WAC1_ACTIVE = 0, 1, 2, 3
FP_KINDS = {"ac1": FP_AC1, "wac1": FP_WAC1, "ac1_active": FP_AC1_ACTIVE, "wac1_active": FP_WAC1_ACTIVE}
MEM_AUTO, MEM_GLOBAL, MEM_STORE_SHARED, MEM_TCN_SHARED, MEM_STORE_CLUSTER = -1, 0, 1, 2, 3
MEM_NAMES = {0: "global", 1: "store_shared", 2: "tcn_shared", 3: "store_cluster"}
NUM_TIMERS = 11
(TIMER_OVERALL, TIMER_PREPROCESSING, TIMER_SEARCH, TIMER_FIXPOINT, TIMER_TRANSFER_CPU2GPU,
 TIMER_TRANSFER_GPU2CPU, TIMER_SELECT_FP_FUNCTIONS, TIMER_WAIT_CPU, TIMER_DIVE,
 TIMER_LATEST_BEST_OBJ_FOUND, TIMER_FIRST_BLOCK_IDLE) = range(NUM_TIMERS)

STATUS_NAMES = {0: "TB_OK", 1: "TB_ERR_INVALID", 2: "TB_ERR_CUDA", 3: "TB_ERR_NOMEM",
                4: "TB_ERR_UNSUPPORTED", 5: "TB_ERR_NO_DEVICE", 6: "TB_ERR_IO", 7: "TB_ERR_PARSE",
                8: "TB_ERR_DEPTH"}

PROP_DTYPE = np.dtype([("op", "<i4"), ("x", "<i4"), ("y", "<i4"), ("z", "<i4")])


class TbProp(C.Structure):
    _fields_ = [("op", C.c_int32), ("x", C.c_int32), ("y", C.c_int32), ("z", C.c_int32)]


class TbStrategy(C.Structure):
    _fields_ = [("var_order", C.c_int32), ("val_order", C.c_int32), ("n", C.c_int32),
                ("vars", C.POINTER(C.c_int32))]


class TbProblem(C.Structure):
    _fields_ = [("nvars", C.c_int32), ("nprops", C.c_int32),
                ("lb", C.POINTER(C.c_int32)), ("ub", C.POINTER(C.c_int32)),
                ("props", C.POINTER(TbProp)),
                ("nstrategies", C.c_int32), ("strategies", C.POINTER(TbStrategy)),
                ("has_eps_strategy", C.c_int32), ("obj_var", C.c_int32)]


class TbOptions(C.Structure):
    _fields_ = [("fixpoint", C.c_int32), ("wac1_threshold", C.c_int32),
                ("subproblems_power", C.c_int32), ("subproblems_factor", C.c_int32),
                ("or_blocks", C.c_int32), ("threads_per_block", C.c_int32),
                ("mem_kind", C.c_int32), ("cluster_size", C.c_int32),
                ("verbose", C.c_int32), ("max_depth", C.c_int32),
                ("gpu_rank", C.c_int32), ("gpu_world", C.c_int32),
                ("device", C.c_int32), ("propagate_repeat", C.c_int32),
                ("timeout_ms", C.c_uint64), ("cutnodes", C.c_uint64), ("seed", C.c_uint64)]


class TbStats(C.Structure):
    _fields_ = [("num_blocks", C.c_int32), ("depth_max", C.c_int32), ("exhaustive", C.c_int32),
                ("threads_per_block", C.c_int32),
                ("mem_kind", C.c_int32), ("cluster_size", C.c_int32),
                ("subproblems_power", C.c_int32), ("blocks_per_sm", C.c_int32),
                ("nodes", C.c_uint64), ("fails", C.c_uint64), ("solutions", C.c_uint64),
                ("eps_num_subproblems", C.c_uint64), ("eps_solved_subproblems", C.c_uint64),
                ("eps_skipped_subproblems", C.c_uint64), ("num_blocks_done", C.c_uint64),
                ("fixpoint_iterations", C.c_uint64), ("num_deductions", C.c_uint64),
                ("bounds_narrowed", C.c_uint64),
                ("shared_bytes", C.c_uint64), ("store_bytes", C.c_uint64), ("prop_bytes", C.c_uint64),
                ("cumulative_time_block_ns", C.c_int64),
                ("timers_ns", C.c_int64 * NUM_TIMERS),
                ("kernel_ms", C.c_double), ("eps_stolen_subproblems", C.c_uint64),
                ("device_bytes", C.c_uint64), ("eps_split_subproblems", C.c_uint64), ("eps_split_parts_solved", C.c_uint64),
                ("fixpoint_in_effect", C.c_int32), ("pad_", C.c_int32)]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if name == "timers_ns" else v
        return d


class TbResultHeader(C.Structure):
    """Wire format of tb_result_pack: this header, then lb[nvars], then ub[nvars] (int32)."""
    _fields_ = [("magic", C.c_uint32), ("nvars", C.c_int32), ("obj_var", C.c_int32), ("has_solution", C.c_int32),
                ("exhaustive", C.c_int32), ("objective", C.c_int32), ("t_best_ns", C.c_int64), ("stats", TbStats)]


RESULT_MAGIC = 0x54425232


class TbDeviceInfo(C.Structure):
    _fields_ = [("cuda_runtime_version", C.c_int32), ("cuda_driver_version", C.c_int32), ("sm_count", C.c_int32),
                ("cc_major", C.c_int32), ("cc_minor", C.c_int32), ("pad_", C.c_int32),
                ("total_global_mem_bytes", C.c_uint64), ("free_global_mem_bytes", C.c_uint64),
                ("stack_limit_bytes", C.c_uint64), ("heap_limit_bytes", C.c_uint64), ("name", C.c_char * 64)]


class TbSimplifyStats(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("iterations", "vars_before", "props_before", "vars_after", "props_after",
                                         "merged_variables", "eliminated_equalities", "eliminated_entailed",
                                         "eliminated_icse", "eliminated_variables", "eliminated_functional", "root_failed")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class TbLayoutInfo(C.Structure):
    _fields_ = [("nclasses", C.c_int32), ("nchunks", C.c_int32), ("nslots", C.c_int32), ("identity", C.c_int32),
                ("class_count", C.c_int32 * 32), ("loads_per_sweep", C.c_uint64), ("wavefronts_per_load", C.c_double)]


def default_options(**kw):
    o = TbOptions()
    o.fixpoint = FP_WAC1
    o.wac1_threshold = 0
    o.subproblems_power = -1
    o.subproblems_factor = 300
    o.mem_kind = MEM_AUTO
    o.gpu_rank, o.gpu_world = 0, 1
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class Problem:
    """Owns the host buffers of one tb_problem (TNF network + strategies + objective)."""

    def __init__(self, lb, ub, props, strategies=None, obj_var=-1, has_eps_strategy=0):
        self.lb = _i32(lb)
        self.ub = _i32(ub)
        if isinstance(props, np.ndarray) and props.dtype == PROP_DTYPE:
            self.props = np.ascontiguousarray(props)
        else:
            p = np.asarray(props, dtype=np.int32).reshape(-1, 4)
            self.props = np.zeros(len(p), dtype=PROP_DTYPE)
            for i, f in enumerate(("op", "x", "y", "z")):
                self.props[f] = p[:, i]
        assert self.lb.shape == self.ub.shape and self.lb.ndim == 1
        if strategies is None:
            strategies = [(VAR_FIRST_FAIL, VAL_MIN, [])]
        self.strategies = [(int(vo), int(va), _i32(vs)) for vo, va, vs in strategies]
        self._strat_arr = (TbStrategy * max(1, len(self.strategies)))()
        for i, (vo, va, vs) in enumerate(self.strategies):
            self._strat_arr[i].var_order = vo
            self._strat_arr[i].val_order = va
            self._strat_arr[i].n = len(vs)
            self._strat_arr[i].vars = _ptr(vs) if len(vs) else None
        self.c = TbProblem()
        self.c.nvars = len(self.lb)
        self.c.nprops = len(self.props)
        self.c.lb = _ptr(self.lb)
        self.c.ub = _ptr(self.ub)
        self.c.props = self.props.ctypes.data_as(C.POINTER(TbProp))
        self.c.nstrategies = len(self.strategies)
        self.c.strategies = self._strat_arr
        self.c.has_eps_strategy = int(has_eps_strategy)
        self.c.obj_var = int(obj_var)

    @property
    def nvars(self):
        return len(self.lb)

    @property
    def nprops(self):
        return len(self.props)

    @property
    def obj_var(self):
        return self.c.obj_var

    @classmethod
    def from_c(cls, cp):
        """Deep-copies a `const tb_problem*` (e.g. from tb_model_problem) into Python-owned buffers."""
        p = cp.contents if hasattr(cp, "contents") else cp
        n, m = p.nvars, p.nprops
        lb = np.ctypeslib.as_array(p.lb, shape=(n,)).copy() if n else np.zeros(0, np.int32)
        ub = np.ctypeslib.as_array(p.ub, shape=(n,)).copy() if n else np.zeros(0, np.int32)
        if m:
            raw = np.ctypeslib.as_array(C.cast(p.props, C.POINTER(C.c_int32)), shape=(m, 4)).copy()
        else:
            raw = np.zeros((0, 4), np.int32)
        strategies = []
        for i in range(p.nstrategies):
            s = p.strategies[i]
            vs = np.ctypeslib.as_array(s.vars, shape=(s.n,)).copy() if s.n else []
            strategies.append((s.var_order, s.val_order, vs))
        return cls(lb, ub, raw, strategies, p.obj_var, p.has_eps_strategy)
