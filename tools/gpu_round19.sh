#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_active.py -m gpu -x -q ) > $O/pytest_r19.log 2>&1
tail -4 $O/pytest_r19.log
B="python bench.py --no-cpu-baseline --no-fixpoint-leg"
run() { echo "== $*" >> $O/exp19.log; timeout 300 "$@" >> $O/exp19.log 2>> $O/exp19.err; }
for w in simplified:trains15 trains15 simplified:example_wordpress7_500; do
  run $B --workload $w --fp wac1_active
done
run env TB_ACTIVE_MIN_CHUNKS=0 $B --workload simplified:accap_a3 --fp wac1_active
python - <<'PY'
import json, sys
sys.path.insert(0, ".")
for line in open("gpurun_out/exp19.log"):
    line = line.strip()
    if line.startswith("{"):
        d = json.loads(line)
        c = d["config"]
        print("   %s tpb %d blocks %d | Gprop/s %.1f nodes/s %.0f fixpoint share %.2f" % (c["memory_configuration"], c["threads_per_block"], c["num_blocks_per_gpu"], d["value"] / 1e9, d["nodes_per_sec"], d["fixpoint_time_share"]))
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
tail -3 $O/exp19.err
