#!/bin/bash
# Round 2, GPU visit 17 (1 GPU): per-node barriers merged + snapshot requested before the decision - parity, then the benches.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_stream.py -m gpu -x -q > $O/v17_tests.txt 2>&1; tail -3 $O/v17_tests.txt
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --strong-ms 0 --no-fixpoint-leg > $O/v17_bench_n1.json 2> $O/v17_bench.err
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --workload simplified:accap_a3 --no-cpu-baseline --strong-ms 0 --no-fixpoint-leg > $O/v17_bench_accap_n1.json 2>> $O/v17_bench.err
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --workload simplified:example_wordpress7_500 --no-cpu-baseline --strong-ms 0 --no-fixpoint-leg > $O/v17_bench_wordpress_n1.json 2>> $O/v17_bench.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/v17_bench*.json")):
    try:
        d = json.load(open(f)); a = d.get("active_set") or {}
        print(f.split("/")[-1], "Gprop/s %.1f nodes/s %.0f e2e %.1f frac %.3f fp_share %.3f | active nodes/s %.0f" % (d["value"] / 1e9, d["nodes_per_sec"], d["e2e"]["value"] / 1e9, d["roofline"]["frac"], d["fixpoint_time_share"], a.get("nodes_per_sec", 0)))
    except Exception as e:
        print(f, "ERR", e)
PY
