"""bench.py on a machine without a GPU: the reference arm (the CPU oracle, the one place besides tests/ and smoke() that
may execute oracle/) prints one JSON line with the contract's keys; the product arm refuses to run (no CPU fallback);
workloads load."""
import json
import os
import subprocess
import sys

import pytest

from turbo_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--workload", "simplified:accap_a3"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "propagations/sec" and d["unit"] == "propagations/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["nodes_per_sec"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload_arg"] == "simplified:accap_a3" and "benchmarks/accap_a3.fzn" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(engine.device_count() > 0, reason="a GPU is present")
def test_product_arm_has_no_cpu_fallback():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_workloads_load():
    sys.path.insert(0, ROOT)
    import bench
    pb, info = bench.load_workload("simplified:trains15")
    full, _ = bench.load_workload("trains15")
    assert 0 < pb.nvars < full.nvars and 0 < pb.nprops < full.nprops and info["objective_kind"] == 0
    assert "simplifier" in bench.data_description("simplified:trains15") and "disable_simplify" in bench.data_description("trains15")
