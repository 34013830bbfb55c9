#!/bin/bash
# Round 2, GPU visit 20 (1 GPU): new automatic placement / block shape (store in shared memory, few-warp blocks) - the whole
# GPU suite, then the three headline networks through bench.py.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/v20_tests.txt 2>&1; tail -4 $O/v20_tests.txt
for w in trains15 accap_a3 example_wordpress7_500; do
  timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --workload simplified:$w --no-cpu-baseline --strong-ms 0 > $O/v20_bench_$w.json 2> $O/v20_bench_$w.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/v20_bench*.json")):
    try:
        d = json.load(open(f)); a = d.get("active_set") or {}; u = d.get("unsimplified_network") or {}; c = d["config"]; k = d.get("fixpoint_kernel") or {}
        print(f.split("/")[-1], c["memory_configuration"], c["num_blocks_per_gpu"], "x", c["threads_per_block"], "Gprop/s %.1f nodes/s %.0f e2e %.1f frac %.3f fp_share %.3f | fixpoint alone %.1f G | active nodes/s %.0f | unsimplified %.1f G %.0f nodes/s" % (
            d["value"] / 1e9, d["nodes_per_sec"], d["e2e"]["value"] / 1e9, d["roofline"]["frac"], d["fixpoint_time_share"], k.get("propagations_per_sec", 0) / 1e9, a.get("nodes_per_sec", 0), u.get("value", 0) / 1e9, u.get("nodes_per_sec", 0)))
    except Exception as e:
        print(f, "ERR", e)
PY
