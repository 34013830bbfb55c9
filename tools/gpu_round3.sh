#!/bin/bash
# GPU visit 3: hybrid table placement — parity, then A/B against TB_NO_HYBRID_TABLE=1.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_known_answers.py -m gpu -x -q ) > $O/pytest_hybrid.log 2>&1
tail -3 $O/pytest_hybrid.log
for w in trains15 simplified:trains15 simplified:example_wordpress7_500; do
  for nh in 0 1; do
    echo "== $w TB_NO_HYBRID_TABLE=$nh" >> $O/exp3.log
    TB_NO_HYBRID_TABLE=$nh timeout 300 python bench.py --workload $w --no-cpu-baseline >> $O/exp3.log 2>> $O/exp3.err
  done
done
python - <<'PY'
import json
for line in open("gpurun_out/exp3.log"):
    line = line.strip()
    if line.startswith("{"):
        d = json.loads(line)
        fk = d.get("fixpoint_kernel", {})
        print("   Gprop/s %.1f nodes/s %.0f frac %.4f | fixpoint kernel Gprop/s %.1f frac %.4f | e2e %.1f" % (
            d["value"] / 1e9, d["nodes_per_sec"], d["roofline"]["frac"], fk.get("propagations_per_sec", 0) / 1e9, fk.get("smem_frac", 0), d["e2e"]["value"] / 1e9))
    else:
        print(line)
PY
