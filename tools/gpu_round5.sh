#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --no-cpu-baseline --no-fixpoint-leg"
run() { echo "== $*" >> $O/exp5.log; timeout 300 $B "$@" >> $O/exp5.log 2>> $O/exp5.err; }
run --workload simplified:trains15
run --workload simplified:trains15 --mem store_shared
run --workload simplified:trains15 --mem store_shared --tpb 1024
run --workload simplified:accap_a3
run --workload simplified:accap_a3 --tpb 128
run --workload simplified:accap_a3 --tpb 512
run --workload simplified:example_wordpress7_500
python - <<'PY'
import json
for line in open("gpurun_out/exp5.log"):
    line = line.strip()
    if line.startswith("{"):
        d = json.loads(line)
        c = d["config"]
        print("   %s tpb %d blocks %d | Gprop/s %.1f nodes/s %.0f frac %.4f fixpoint share %.2f" % (
            c["memory_configuration"], c["threads_per_block"], c["num_blocks_per_gpu"], d["value"] / 1e9, d["nodes_per_sec"], d["roofline"]["frac"], d["fixpoint_time_share"]))
    else:
        print(line)
PY
