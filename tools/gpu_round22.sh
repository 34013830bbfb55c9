#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/pytest_gpu_final2.log 2>&1
tail -6 $O/pytest_gpu_final2.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
B="python bench.py --no-cpu-baseline --no-fixpoint-leg"
for w in simplified:trains15 trains15 simplified:example_wordpress7_500; do
  echo "== $w wac1_active" >> $O/exp22.log
  timeout 300 $B --workload $w --fp wac1_active >> $O/exp22.log 2>> $O/exp22.err
done
python - <<'PY'
import json
for line in open("gpurun_out/exp22.log"):
    line = line.strip()
    if line.startswith("{"):
        d = json.loads(line); c = d["config"]
        print("   %s tpb %d blocks %d | Gprop/s %.1f nodes/s %.0f" % (c["memory_configuration"], c["threads_per_block"], c["num_blocks_per_gpu"], d["value"] / 1e9, d["nodes_per_sec"]))
    else:
        print(line)
PY
