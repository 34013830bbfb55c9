#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_active.py -x -q ) > $O/pytest_active2.log 2>&1
tail -6 $O/pytest_active2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -f -o $O/solve_active_strains \
  python bench.py --workload simplified:trains15 --fp wac1_active --steps 1 --warmup 0 --cutnodes 300 --no-cpu-baseline --no-fixpoint-leg > $O/ncu_solve_active.log 2>&1
tail -2 $O/ncu_solve_active.log
python - <<'PY'
import sys
sys.path.insert(0, ".")
from tests import golden_io
from turbo_b200 import abi, engine
pb, info = golden_io.load_simplified_problem("trains15")
for fp in (abi.FP_WAC1, abi.FP_WAC1_ACTIVE):
    with engine.Solver(pb, cutnodes=2000, fixpoint=fp) as s:
        r = s.solve()
    st = r["stats"]
    print("fp", fp, "nodes", st["nodes"], "sweeps/node %.2f" % (st["fixpoint_iterations"] / st["nodes"]), "evals/node %.0f" % (st["num_deductions"] / st["nodes"]),
          "fails", st["fails"], "kernel_ms %.1f" % st["kernel_ms"])
PY
