#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_active.py -x -q ) > $O/pytest_active.log 2>&1
tail -25 $O/pytest_active.log
B="python bench.py --no-cpu-baseline --no-fixpoint-leg"
run() { echo "== $*" >> $O/exp6.log; timeout 300 $B "$@" >> $O/exp6.log 2>> $O/exp6.err; }
run --workload trains15
run --workload trains15 --fp wac1_active
run --workload simplified:trains15
run --workload simplified:trains15 --fp wac1_active
run --workload simplified:example_wordpress7_500 --fp wac1_active
run --workload simplified:accap_a3
run --workload simplified:accap_a3 --fp wac1_active
python - <<'PY'
import json
for line in open("gpurun_out/exp6.log"):
    line = line.strip()
    if line.startswith("{"):
        d = json.loads(line)
        c = d["config"]
        print("   %s tpb %d blocks %d | Gprop/s %.1f nodes/s %.0f frac %.4f fixpoint share %.2f" % (
            c["memory_configuration"], c["threads_per_block"], c["num_blocks_per_gpu"], d["value"] / 1e9, d["nodes_per_sec"], d["roofline"]["frac"], d["fixpoint_time_share"]))
    else:
        print(line)
PY
tail -5 $O/exp6.err
