"""The C ABI boundary: the shared library loads without a GPU, exports every symbol that
include/turbo_b200.h declares, validates its inputs, and fails loudly (never falls back to a CPU
path) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from tests import tnf_gen
from turbo_b200 import abi, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "turbo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    # (a name followed by "(*" is the return type of a function-pointer typedef, not a function)
    return sorted(set(re.findall(r"\b(tb_[a-z0-9_]+)\s*\((?!\s*\*)", text)))


def test_library_exports_every_declared_symbol():
    L = engine.lib()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} is declared in include/turbo_b200.h but not exported"


def test_struct_layouts_match_the_header():
    assert C.sizeof(abi.TbProp) == 16
    assert C.sizeof(abi.TbStrategy) == 24
    assert C.sizeof(abi.TbProblem) == 56
    assert C.sizeof(abi.TbOptions) == 80
    assert C.sizeof(abi.TbStats) == 32 + 8 * 10 + 8 * 3 + 8 + 8 * abi.NUM_TIMERS + 8 + 8 + 8 + 8 + 16


def test_version_and_device_count():
    assert b"sm_100a" in engine.lib().tb_version()
    assert engine.device_count() >= 0


def test_invalid_problems_are_rejected():
    pb = tnf_gen.planted(10, 10, 0)
    bad = abi.Problem(pb.lb, pb.ub, np.array([[abi.OP_ADD, 0, 1, 99]], np.int32))
    with pytest.raises(engine.TurboError) as e:
        engine.Solver(bad)
    assert e.value.status == 1
    bad = abi.Problem(pb.lb, pb.ub, np.array([[42, 0, 1, 2]], np.int32))
    with pytest.raises(engine.TurboError):
        engine.Solver(bad)
    with pytest.raises(engine.TurboError):
        engine.Solver(pb, gpu_rank=3, gpu_world=2)


@pytest.mark.skipif(engine.device_count() > 0, reason="a GPU is present")
def test_no_gpu_means_loud_failure_not_a_cpu_fallback():
    pb = tnf_gen.planted(10, 10, 0)
    with pytest.raises(engine.TurboError) as e:
        engine.Solver(pb)
    assert e.value.status == 5 and "no CPU fallback" in str(e.value)
    exe = os.path.join(ROOT, "turbo_b200", "bin", "turbo")
    fzn = os.path.join(ROOT, "tests", "data", "tiny.fzn")
    r = subprocess.run([exe, "-s", fzn], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    r = subprocess.run([exe, "-arch", "cpu", fzn], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_cli_usage_and_parse_errors():
    exe = os.path.join(ROOT, "turbo_b200", "bin", "turbo")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0 and r.stdout.startswith("usage:")
    r = subprocess.run([exe, "-or", "2", "-p", "2", "x.fzn"], capture_output=True, text=True)
    assert r.returncode != 0 and "cannot be used at the same time" in r.stderr
    r = subprocess.run([exe, "/nonexistent.fzn"], capture_output=True, text=True)
    assert r.returncode != 0 and "Could not parse input file." in r.stderr
    r = subprocess.run([exe, "-eps_var_order", "first_fail", os.path.join(ROOT, "tests", "data", "tiny.fzn")], capture_output=True, text=True)
    assert r.returncode != 0 and "must be specified together" in r.stdout


def test_product_never_links_the_oracle():
    out = subprocess.run(["ldd", os.path.join(ROOT, "turbo_b200", "libturbo_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out
    for dirpath, _, files in os.walk(os.path.join(ROOT, "turbo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "tnf_oracle" not in text and "oracle_py" not in text, os.path.join(dirpath, f)


def test_minizinc_solver_entry():
    """The .msc MiniZinc reads (reference benchmarks/minizinc/turbo.gpu.release.msc): valid JSON, points at the built
    driver and at a solver library that compiles set variables away, advertises the flags the driver parses."""
    import json
    d = os.path.join(ROOT, "turbo_b200", "minizinc")
    msc = json.load(open(os.path.join(d, "turbo.b200.release.msc")))
    assert os.path.exists(os.path.normpath(os.path.join(d, msc["executable"])))
    assert "nosets.mzn" in open(os.path.join(d, msc["mznlib"], "redefinitions.mzn")).read()
    assert msc["supportsFzn"] and not msc["supportsMzn"] and set(["-a", "-n", "-p", "-s", "-v", "-f", "-t"]) <= set(msc["stdFlags"])
    exe = os.path.join(ROOT, "turbo_b200", "bin", "turbo")
    usage = subprocess.run([exe], capture_output=True, text=True).stdout
    for flag, *_ in msc["extraFlags"]:
        assert flag in usage or flag in ("-timeout",), flag
