#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_active.py -m gpu -x -q ) > $O/pytest_active4.log 2>&1
tail -5 $O/pytest_active4.log
B="python bench.py --no-cpu-baseline --no-fixpoint-leg"
run() { echo "== $*" >> $O/exp12.log; timeout 300 "$@" >> $O/exp12.log 2>> $O/exp12.err; }
for w in simplified:trains15 trains15 simplified:example_wordpress7_500 simplified:accap_a3; do
  run $B --workload $w --fp wac1_active
done
run $B --workload simplified:trains15 --fp wac1_active --mem tcn_shared
python - <<'PY'
import json, sys, os
sys.path.insert(0, ".")
for line in open("gpurun_out/exp12.log"):
    line = line.strip()
    if line.startswith("{"):
        d = json.loads(line)
        c = d["config"]
        print("   %s tpb %d blocks %d | Gprop/s %.1f nodes/s %.0f frac %.4f fixpoint share %.2f" % (
            c["memory_configuration"], c["threads_per_block"], c["num_blocks_per_gpu"], d["value"] / 1e9, d["nodes_per_sec"], d["roofline"]["frac"], d["fixpoint_time_share"]))
    else:
        print(line)
from tests import golden_io
from turbo_b200 import abi, engine
pb, info = golden_io.load_simplified_problem("trains15")
for fp in (abi.FP_WAC1, abi.FP_WAC1_ACTIVE):
    with engine.Solver(pb, cutnodes=2000, fixpoint=fp) as s:
        r = s.solve()
    st = r["stats"]
    print("fp", fp, "nodes", st["nodes"], "sweeps/node %.2f" % (st["fixpoint_iterations"] / st["nodes"]), "evals/node %.0f" % (st["num_deductions"] / st["nodes"]),
          "fails", st["fails"], "kernel_ms %.1f" % st["kernel_ms"], "obj", r["objective"])
PY
tail -3 $O/exp12.err
