"""The `turbo` driver end to end on a GPU: the reference's process surface (src/turbo.cpp, include/statistics.hpp:338-412,
include/config.hpp:237-266) — flags, the %%%mzn-stat protocol, the solution block, the final separator — with and without
the TNF simplifier.  Inputs: tests/data/tiny.fzn and golden networks written as .tnf files (the FlatZinc sources of the
reference are not on the GPU box)."""
import os
import re
import subprocess

import pytest

from tests import golden_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "turbo_b200", "bin", "turbo")
TINY = os.path.join(ROOT, "tests", "data", "tiny.fzn")
pytestmark = pytest.mark.gpu


def run(*args, timeout=300):
    r = subprocess.run([EXE, *args], capture_output=True, text=True, timeout=timeout)
    stats = dict(re.findall(r"^%%%mzn-stat: (\w+)=(.*)$", r.stdout, flags=re.M))
    return r, stats


@pytest.mark.parametrize("extra", [[], ["-disable_simplify"], ["-fp", "ac1"], ["-or", "4", "-sub", "3"], ["-globalmem"]])
def test_tiny_model_protocol(extra):
    r, st = run("-s", "-v", *extra, TINY)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    assert out.startswith('%%%mzn-stat: command_line="')
    # optimum: a + b + c >= 6, a != b, c >= 3, minimise 3a + 5b + 2c  ->  a=2, b=0, c=4 (cost 14)
    assert "----------" in out and "==========" in out
    sol = dict(re.findall(r"^(\w+) = (.*);$", out, flags=re.M))
    a, b, c, cost = int(sol["a"]), int(sol["b"]), int(sol["c"]), int(sol["cost"])
    assert 3 * a + 5 * b + 2 * c == cost and a + b + c >= 6 and a != b and c >= 3 and sol["big"] == "true"
    assert sol["items"] == f"array1d(1..3, [{a}, {b}, {c}])"
    assert int(st["objective"]) == cost == 14
    for key in ("parsed_variables", "tcn_variables", "tcn_constraints", "num_blocks", "memory_configuration", "nodes", "failures",
                "variables", "propagators", "peakDepth", "initTime", "solveTime", "num_solutions", "eps_num_subproblems",
                "fixpoint_iterations", "num_deductions", "solve_time", "search_time", "fixpoint_time", "best_obj_time"):
        assert key in st, key
    if "-disable_simplify" in extra:
        assert "preprocessed_tcn_variables" not in st and st["variables"] == st["tcn_variables"]
    else:
        assert int(st["preprocessed_tcn_variables"]) <= int(st["tcn_variables"])
        assert st["variables"] == st["preprocessed_tcn_variables"] and st["propagators"] == st["preprocessed_tcn_constraints"]


@pytest.mark.parametrize("name", ["pat2", "pennies5", "sudoku_opt3", "maximize_unconstrained", "bug2", "accap_a3"])
@pytest.mark.parametrize("simplify", [True, False])
def test_golden_networks_through_the_driver(tmp_path, name, simplify):
    pb, info = golden_io.load(name)
    path = str(tmp_path / (name + ".tnf"))
    golden_io.write_tnf(path, pb, info)
    args = ["-s"] + ([] if simplify else ["-disable_simplify"])
    if info["expected"] is None:
        args += ["-t", "3000"]
    r, st = run(*args, path)
    assert r.returncode == 0, r.stderr          # exit code 2 = the solution failed the re-check against the full network
    assert "----------" in r.stdout
    if info["expected"] is not None:
        assert "==========" in r.stdout and int(st["objective"]) == info["expected"]
    else:
        assert "objective" in st
    if simplify:
        assert int(st["preprocessed_tcn_variables"]) < pb.nvars or pb.nvars <= 8


def test_unsatisfiable_and_timeout(tmp_path):
    p = tmp_path / "unsat.fzn"
    p.write_text("var 0..5: x;\nvar 0..5: y;\nconstraint int_lin_eq([1,1],[x,y],20);\nsolve satisfy;\n")
    r, _ = run("-s", str(p))
    assert r.returncode == 0 and "=====UNSATISFIABLE=====" in r.stdout
    pb, info = golden_io.load("trains15")
    path = str(tmp_path / "trains15.tnf")
    golden_io.write_tnf(path, pb, info)
    r, st = run("-s", "-t", "2000", path)
    assert r.returncode == 0 and "==========" not in r.stdout
    assert float(st["solveTime"]) < 10.0
