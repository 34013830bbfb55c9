"""ctypes binding of the host front-end (tb_model_* in include/turbo_b200.h): FlatZinc -> TNF,
synthetic networks, .tnf files, solution printing and checking. Runs without a GPU."""
import ctypes as C

import numpy as np

from . import abi
from .engine import TurboError, lib

_INIT = False
FIXPOINT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(abi.TbProblem), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                          C.POINTER(C.c_int32))


def _lib():
    global _INIT
    L = lib()
    if not _INIT:
        vp, i32p = C.c_void_p, C.POINTER(C.c_int32)
        L.tb_model_load_fzn.argtypes = [C.POINTER(vp), C.c_char_p, C.c_uint32]
        L.tb_model_parse_fzn.argtypes = [C.POINTER(vp), C.c_char_p, C.c_size_t, C.c_uint32]
        L.tb_model_synthetic.argtypes = [C.POINTER(vp), C.c_int32, C.c_int32, C.c_uint64]
        L.tb_model_load_tnf.argtypes = [C.POINTER(vp), C.c_char_p]
        L.tb_model_save_tnf.argtypes = [vp, C.c_char_p]
        L.tb_model_push_eps_strategy.argtypes = [vp, C.c_int32, C.c_int32]
        L.tb_model_problem.argtypes = [vp]
        L.tb_model_problem.restype = C.POINTER(abi.TbProblem)
        for n in ("tb_model_objective_kind", "tb_model_user_objective_var", "tb_model_num_parsed_variables",
                  "tb_model_num_parsed_constraints", "tb_model_root_failed"):
            getattr(L, n).argtypes = [vp]
            getattr(L, n).restype = C.c_int32
        L.tb_model_check_solution.argtypes = [vp, i32p, i32p]
        L.tb_model_check_solution.restype = C.c_int32
        L.tb_model_check_tnf.argtypes = [vp, i32p]
        L.tb_model_check_tnf.restype = C.c_int32
        L.tb_model_format_solution.argtypes = [vp, i32p, i32p, C.c_char_p, C.c_size_t]
        L.tb_model_format_solution.restype = C.c_size_t
        L.tb_model_simplify.argtypes = [vp, vp, vp, C.POINTER(abi.TbSimplifyStats)]
        L.tb_model_simplify.restype = C.c_int
        L.tb_model_num_full_variables.argtypes = [vp]
        L.tb_model_num_full_variables.restype = C.c_int32
        L.tb_model_expand_solution.argtypes = [vp, i32p, i32p, i32p, i32p]
        L.tb_model_expand_solution.restype = C.c_int
        L.tb_model_destroy.argtypes = [vp]
        L.tb_model_destroy.restype = None
        _INIT = True
    return L


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class Model:
    def __init__(self, handle):
        self._h = handle
        self._refresh()

    def _refresh(self):
        L = _lib()
        self.problem = abi.Problem.from_c(L.tb_model_problem(self._h))
        self.objective_kind = L.tb_model_objective_kind(self._h)      # -1 satisfy, 0 min, 1 max
        self.user_objective_var = L.tb_model_user_objective_var(self._h)
        self.root_failed = bool(L.tb_model_root_failed(self._h))
        self.parsed_variables = L.tb_model_num_parsed_variables(self._h)
        self.parsed_constraints = L.tb_model_num_parsed_constraints(self._h)

    @staticmethod
    def _make(fn, *args):
        h = C.c_void_p()
        rc = fn(C.byref(h), *args)
        if rc != 0:
            raise TurboError(rc, lib().tb_last_error().decode())
        return Model(h)

    @classmethod
    def from_fzn(cls, path, flags=0):
        return cls._make(_lib().tb_model_load_fzn, str(path).encode(), flags)

    @classmethod
    def from_fzn_text(cls, text, flags=0):
        b = text.encode()
        return cls._make(_lib().tb_model_parse_fzn, b, len(b), flags)

    @classmethod
    def synthetic(cls, nvars, nprops, seed=0xB200):
        return cls._make(_lib().tb_model_synthetic, nvars, nprops, seed)

    @classmethod
    def from_tnf(cls, path):
        return cls._make(_lib().tb_model_load_tnf, str(path).encode())

    def save_tnf(self, path):
        rc = _lib().tb_model_save_tnf(self._h, str(path).encode())
        if rc != 0:
            raise TurboError(rc, lib().tb_last_error().decode())

    def push_eps_strategy(self, var_order, val_order):
        _lib().tb_model_push_eps_strategy(self._h, var_order, val_order)
        self._refresh()

    def simplify(self, fixpoint="device", device=0):
        """TNF simplifier (tb_model_simplify). `fixpoint` is "device" (the engine's tb_propagate on a GPU, what
        the `turbo` driver uses) or a callable (problem: abi.Problem) -> (lb, ub, failed) — tests pass the oracle.
        Afterwards `self.problem` is the reduced network."""
        L = _lib()
        st = abi.TbSimplifyStats()
        if fixpoint == "device":
            dev = C.c_int32(device)
            fn = C.cast(L.tb_fixpoint_on_device, C.c_void_p)
            rc = L.tb_model_simplify(self._h, fn, C.cast(C.byref(dev), C.c_void_p), C.byref(st))
        else:
            i32p = C.POINTER(C.c_int32)
            err = []

            @FIXPOINT_FN
            def cb(ctx, pbp, lbp, ubp, failedp):
                try:
                    pb = abi.Problem.from_c(pbp)
                    n = pb.nvars
                    lb = np.ctypeslib.as_array(lbp, shape=(max(1, n),))
                    ub = np.ctypeslib.as_array(ubp, shape=(max(1, n),))
                    nlb, nub, failed = fixpoint(pb)
                    failedp[0] = 1 if failed else 0
                    if not failed and n:
                        lb[:n] = nlb[:n]
                        ub[:n] = nub[:n]
                    return 0
                except Exception as e:      # never unwind through C
                    err.append(e)
                    return abi.ERR_INVALID if hasattr(abi, "ERR_INVALID") else 1
            rc = L.tb_model_simplify(self._h, C.cast(cb, C.c_void_p), None, C.byref(st))
            if err:
                raise err[0]
        if rc != 0:
            raise TurboError(rc, lib().tb_last_error().decode())
        self._refresh()
        self.simplify_stats = st.as_dict()
        return self.simplify_stats

    @property
    def num_full_variables(self):
        return int(_lib().tb_model_num_full_variables(self._h))

    def expand(self, lb, ub=None):
        """Store of `self.problem` (possibly reduced) -> store of the network as it was built."""
        lb = np.ascontiguousarray(lb, dtype=np.int32)
        ub = lb if ub is None else np.ascontiguousarray(ub, dtype=np.int32)
        n = max(1, self.num_full_variables)
        flb, fub = np.zeros(n, np.int32), np.zeros(n, np.int32)
        rc = _lib().tb_model_expand_solution(self._h, _p(lb), _p(ub), _p(flb), _p(fub))
        if rc != 0:
            raise TurboError(rc, lib().tb_last_error().decode())
        return flb[:self.num_full_variables], fub[:self.num_full_variables]

    def check_solution(self, lb, ub=None):
        """Violated FlatZinc constraints at the point lb (0 = valid; -1 = no FlatZinc source)."""
        lb = np.ascontiguousarray(lb, dtype=np.int32)
        ub = lb if ub is None else np.ascontiguousarray(ub, dtype=np.int32)
        return int(_lib().tb_model_check_solution(self._h, _p(lb), _p(ub)))

    def check_tnf(self, lb):
        lb = np.ascontiguousarray(lb, dtype=np.int32)
        return int(_lib().tb_model_check_tnf(self._h, _p(lb)))

    def format_solution(self, lb, ub=None):
        lb = np.ascontiguousarray(lb, dtype=np.int32)
        ub = lb if ub is None else np.ascontiguousarray(ub, dtype=np.int32)
        n = _lib().tb_model_format_solution(self._h, _p(lb), _p(ub), None, 0)
        buf = C.create_string_buffer(n + 1)
        _lib().tb_model_format_solution(self._h, _p(lb), _p(ub), buf, n + 1)
        return buf.value.decode()

    def user_objective(self, lb, ub):
        """The objective value as the reference prints it: lb for minimise, ub of the original
        variable for maximise (statistics.hpp:378-388)."""
        if self.objective_kind < 0:
            return None
        v = self.user_objective_var
        return int(lb[v]) if self.objective_kind == 0 else int(ub[v])

    def close(self):
        if self._h:
            _lib().tb_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
