"""ctypes binding of turbo_b200/libturbo_b200.so (the product's C ABI, include/turbo_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is visible when a
solver is created, this raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# TURBO_B200_LIB selects another build of the same library (A/B runs of kernel variants); there is still no fallback.
LIB_PATH = os.environ.get("TURBO_B200_LIB") or os.path.join(_HERE, "libturbo_b200.so")
_LIB = None


class TurboError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{abi.STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C turbo_b200`). The engine has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        i32p = C.POINTER(C.c_int32)
        vp = C.c_void_p
        L.tb_create.argtypes = [C.POINTER(vp), C.POINTER(abi.TbProblem), C.POINTER(abi.TbOptions)]
        L.tb_propagate.argtypes = [vp, i32p, i32p, i32p, i32p, i32p, C.POINTER(abi.TbStats)]
        L.tb_propagate_batch.argtypes = [vp, C.c_int32, i32p, i32p, i32p, i32p, i32p, C.POINTER(abi.TbStats)]
        L.tb_dive.argtypes = [vp, C.c_uint64, C.c_int32, i32p, i32p, i32p, i32p]
        L.tb_dive_batch.argtypes = [vp, C.c_uint64, C.c_int32, C.c_int32, i32p, i32p, i32p, i32p]
        L.tb_solve.argtypes = [vp, i32p, i32p, i32p, i32p, i32p, C.POINTER(abi.TbStats)]
        L.tb_link_peers.argtypes = [C.POINTER(vp), C.c_int32]
        L.tb_export_bound_handle.argtypes = [vp, vp]
        L.tb_import_peer_bounds.argtypes = [vp, vp, C.c_int32]
        L.tb_read_bound.argtypes = [vp, i32p]
        L.tb_get_config.argtypes = [vp, C.POINTER(abi.TbStats)]
        L.tb_layout_describe.argtypes = [C.POINTER(abi.TbProblem), C.c_int32, C.POINTER(abi.TbLayoutInfo), i32p]
        L.tb_layout_class_name.argtypes = [C.c_int32]
        L.tb_layout_class_name.restype = C.c_char_p
        L.tb_destroy.argtypes = [vp]
        L.tb_destroy.restype = None
        L.tb_last_error.restype = C.c_char_p
        L.tb_version.restype = C.c_char_p
        L.tb_device_count.restype = C.c_int32
        for name in ("tb_create", "tb_propagate", "tb_propagate_batch", "tb_dive", "tb_dive_batch", "tb_solve",
                     "tb_link_peers", "tb_export_bound_handle", "tb_import_peer_bounds", "tb_read_bound",
                     "tb_get_config"):
            getattr(L, name).restype = C.c_int
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise TurboError(rc, lib().tb_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def layout_describe(problem, nbanks=16):
    """The host layout pass on its own (no GPU): class histogram, variable placement, bank model."""
    L = lib()
    info = abi.TbLayoutInfo()
    slot_of = np.zeros(max(1, problem.nvars), dtype=np.int32)
    _check(L.tb_layout_describe(C.byref(problem.c), nbanks, C.byref(info), _p(slot_of)))
    classes = {L.tb_layout_class_name(c).decode(): info.class_count[c] for c in range(info.nclasses)}
    return dict(classes=classes, nchunks=info.nchunks, nslots=info.nslots, identity=bool(info.identity),
                loads_per_sweep=info.loads_per_sweep, wavefronts_per_load=info.wavefronts_per_load,
                slot_of=slot_of[:problem.nvars])


def layout_watch_lists(problem, nbanks=16):
    """(off, list, slot_of, chunk_of_prop) of the active-set fixpoint's watch lists (host only, no GPU)."""
    L = lib()
    i32p = C.POINTER(C.c_int32)
    L.tb_layout_watch_lists.argtypes = [C.POINTER(abi.TbProblem), C.c_int32, i32p, i32p, i32p, i32p, i32p, i32p]
    L.tb_layout_watch_lists.restype = C.c_int
    ns, ne = C.c_int32(0), C.c_int32(0)
    _check(L.tb_layout_watch_lists(C.byref(problem.c), nbanks, C.byref(ns), C.byref(ne), None, None, None, None))
    off = np.zeros(ns.value + 1, np.int32)
    lst = np.zeros(max(1, ne.value), np.int32)
    slot_of = np.zeros(max(1, problem.nvars), np.int32)
    chunk_of = np.zeros(max(1, problem.nprops), np.int32)
    _check(L.tb_layout_watch_lists(C.byref(problem.c), nbanks, C.byref(ns), C.byref(ne), _p(off), _p(lst), _p(slot_of), _p(chunk_of)))
    return off, lst[:ne.value], slot_of[:problem.nvars], chunk_of[:problem.nprops]


def device_count():
    return int(lib().tb_device_count())


def result_reduce(buffers):
    """tb_result_reduce over a list of packed results (bytes / uint8 arrays of equal length), rank order."""
    L = lib()
    L.tb_result_reduce.argtypes = [C.c_void_p, C.c_int32, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(abi.TbStats), C.POINTER(C.c_int32)]
    L.tb_result_reduce.restype = C.c_int
    bufs = [np.frombuffer(bytes(b), dtype=np.uint8) for b in buffers]
    stride = len(bufs[0])
    flat = np.ascontiguousarray(np.concatenate(bufs))
    hdr = abi.TbResultHeader.from_buffer_copy(flat[:C.sizeof(abi.TbResultHeader)].tobytes())
    nv = max(1, hdr.nvars)
    lb, ub = np.zeros(nv, np.int32), np.zeros(nv, np.int32)
    has, exh, best = C.c_int32(0), C.c_int32(0), C.c_int32(-1)
    st = abi.TbStats()
    _check(L.tb_result_reduce(flat.ctypes.data_as(C.c_void_p), len(bufs), stride, _p(lb), _p(ub), C.byref(has), C.byref(exh),
                              C.byref(st), C.byref(best)))
    return dict(lb=lb[:hdr.nvars], ub=ub[:hdr.nvars], has_solution=bool(has.value), exhaustive=bool(exh.value),
                stats=st.as_dict(), best_rank=best.value)


def device_info(device=0):
    L = lib()
    L.tb_get_device_info.argtypes = [C.c_int32, C.POINTER(abi.TbDeviceInfo)]
    L.tb_get_device_info.restype = C.c_int
    info = abi.TbDeviceInfo()
    _check(L.tb_get_device_info(device, C.byref(info)))
    d = {n: getattr(info, n) for n, _ in info._fields_ if n not in ("pad_", "name")}
    d["name"] = info.name.decode()
    return d


def measure_smem_peak(device=0):
    """Measured shared-memory read bandwidth (GB/s over all SMs) and bytes/clk/SM at the maximum SM clock."""
    L = lib()
    L.tb_measure_smem_peak.argtypes = [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.tb_measure_smem_peak.restype = C.c_int
    gbs, bpc = C.c_double(0), C.c_double(0)
    _check(L.tb_measure_smem_peak(device, C.byref(gbs), C.byref(bpc)))
    return dict(gb_per_s=gbs.value, bytes_per_clk_per_sm=bpc.value)


class Solver:
    """One engine instance bound to one CUDA device (tb_solver)."""

    def __init__(self, problem, **options):
        self.problem = problem
        self.options = abi.default_options(**options)
        self._h = C.c_void_p()
        _check(lib().tb_create(C.byref(self._h), C.byref(problem.c), C.byref(self.options)))

    def close(self):
        if self._h:
            lib().tb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def handle(self):
        return self._h

    def config(self):
        st = abi.TbStats()
        _check(lib().tb_get_config(self._h, C.byref(st)))
        return st.as_dict()

    def propagate(self, lb=None, ub=None):
        n = self.problem.nvars
        if lb is not None:
            lb = np.ascontiguousarray(lb, dtype=np.int32)
            ub = np.ascontiguousarray(ub, dtype=np.int32)
        olb = np.zeros(max(1, n), np.int32)
        oub = np.zeros(max(1, n), np.int32)
        failed = C.c_int32(0)
        st = abi.TbStats()
        _check(lib().tb_propagate(self._h, _p(lb), _p(ub), _p(olb), _p(oub), C.byref(failed), C.byref(st)))
        return dict(lb=olb[:n], ub=oub[:n], failed=bool(failed.value), stats=st.as_dict())

    def propagate_batch(self, lb, ub):
        lb = np.ascontiguousarray(lb, dtype=np.int32)
        ub = np.ascontiguousarray(ub, dtype=np.int32)
        b = lb.shape[0]
        olb = np.zeros_like(lb)
        oub = np.zeros_like(ub)
        failed = np.zeros(b, np.int32)
        st = abi.TbStats()
        _check(lib().tb_propagate_batch(self._h, b, _p(lb), _p(ub), _p(olb), _p(oub), _p(failed), C.byref(st)))
        return dict(lb=olb, ub=oub, failed=failed.astype(bool), stats=st.as_dict())

    def dive_batch(self, first, count, depth):
        n = self.problem.nvars
        olb = np.zeros((count, max(1, n)), np.int32)
        oub = np.zeros((count, max(1, n)), np.int32)
        if n:
            olb = np.zeros((count, n), np.int32)
            oub = np.zeros((count, n), np.int32)
        rem = np.zeros(count, np.int32)
        kind = np.zeros(count, np.int32)
        _check(lib().tb_dive_batch(self._h, first, count, depth, _p(olb), _p(oub), _p(rem), _p(kind)))
        return dict(lb=olb, ub=oub, remaining_depth=rem, leaf_kind=kind)

    def dive(self, idx, depth):
        r = self.dive_batch(idx, 1, depth)
        return dict(lb=r["lb"][0], ub=r["ub"][0], remaining_depth=int(r["remaining_depth"][0]),
                    leaf_kind=int(r["leaf_kind"][0]))

    def solve(self, stop_flag=None):
        n = max(1, self.problem.nvars)
        lb = np.zeros(n, np.int32)
        ub = np.zeros(n, np.int32)
        has = C.c_int32(0)
        exh = C.c_int32(0)
        st = abi.TbStats()
        _check(lib().tb_solve(self._h, stop_flag, _p(lb), _p(ub), C.byref(has), C.byref(exh), C.byref(st)))
        obj = None
        if has.value and self.problem.obj_var >= 0:
            obj = int(lb[self.problem.obj_var])
        return dict(lb=lb[:self.problem.nvars], ub=ub[:self.problem.nvars], has_solution=bool(has.value),
                    exhaustive=bool(exh.value), objective=obj, stats=st.as_dict())

    def export_bound_handle(self):
        buf = C.create_string_buffer(64)
        _check(lib().tb_export_bound_handle(self._h, C.cast(buf, C.c_void_p)))
        return buf.raw

    def import_peer_bounds(self, handles):
        blob = b"".join(handles)
        buf = C.create_string_buffer(blob, len(blob))
        _check(lib().tb_import_peer_bounds(self._h, C.cast(buf, C.c_void_p), len(handles)))

    def stream_solutions(self, slots=16):
        """-i / -a: improving solutions also go to a ring the host can read while solve() runs (tb_stream_solutions)."""
        L = lib()
        L.tb_stream_solutions.argtypes = [C.c_void_p, C.c_int32]
        L.tb_stream_solutions.restype = C.c_int
        _check(L.tb_stream_solutions(self._h, slots))

    def poll_solution(self):
        """The newest streamed solution not returned yet, or None (call from another thread while solve() blocks)."""
        L = lib()
        i32p = C.POINTER(C.c_int32)
        L.tb_poll_solution.argtypes = [C.c_void_p, i32p, i32p, i32p, C.POINTER(C.c_int64)]
        L.tb_poll_solution.restype = C.c_int32
        n = self.problem.nvars
        lb, ub = np.zeros(max(1, n), np.int32), np.zeros(max(1, n), np.int32)
        obj, t = C.c_int32(0), C.c_int64(0)
        rc = L.tb_poll_solution(self._h, _p(lb), _p(ub), C.byref(obj), C.byref(t))
        if rc < 0:
            _check(-rc)
        return None if rc == 0 else dict(lb=lb[:n], ub=ub[:n], objective=obj.value, time_ns=t.value)

    def set_timeout(self, ms):
        L = lib()
        L.tb_set_timeout.argtypes = [C.c_void_p, C.c_uint64]
        L.tb_set_timeout.restype = C.c_int
        _check(L.tb_set_timeout(self._h, ms))

    def result_pack(self):
        """The latest solve() of this solver as bytes (tb_result_pack): what one rank contributes to the final gather."""
        L = lib()
        L.tb_result_size.argtypes = [C.c_void_p]
        L.tb_result_size.restype = C.c_size_t
        L.tb_result_pack.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.tb_result_pack.restype = C.c_int
        n = L.tb_result_size(self._h)
        buf = np.zeros(n, np.uint8)
        _check(L.tb_result_pack(self._h, buf.ctypes.data_as(C.c_void_p), n))
        return buf

    def read_bound(self):
        v = C.c_int32(0)
        _check(lib().tb_read_bound(self._h, C.byref(v)))
        return v.value


def link_peers(solvers):
    arr = (C.c_void_p * len(solvers))(*[s.handle for s in solvers])
    _check(lib().tb_link_peers(arr, len(solvers)))
