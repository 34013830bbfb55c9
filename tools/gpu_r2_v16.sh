#!/bin/bash
# Round 2, GPU visit 16 (8 GPUs): the shared tail at N = 8 - capped-objective accap_a3 (fixed regime) on 1 / 8 GPUs with and
# without sharing, the accap_a3 fixed-pool strong-scaling leg, and the driver with -gpus 8 (accap_a3 45 s, trains15 10 s).
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
TAIL_BUDGET_MS=8000 TAIL_ONLY_CAPPED=1 timeout 300 python tools/tail_share.py $N 2>&1 | tee $O/v16_tail_share_n$N.jsonl
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus $N --steps 3 --warmup 3 --workload simplified:accap_a3 --strong-sub 24 --strong-ms 4000 --no-cpu-baseline --e2e-steps 3 > $O/v16_bench_accap_n$N.json 2> $O/v16_bench.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/v16_bench*.json")):
    try:
        d = json.load(open(f)); s = d.get("strong_scaling") or {}
        print(f.split("/")[-1], "n_gpus", d["n_gpus"], "Gprop/s %.1f nodes/s %.0f | strong: nodes/s %.0f solved %s stolen %s best %s" % (
            d["value"] / 1e9, d["nodes_per_sec"], s.get("nodes_per_sec", 0), s.get("subproblems_solved"), s.get("subproblems_stolen"), s.get("best_objective")))
    except Exception as e:
        print(f, "ERR", e)
PY
python - <<PY
import sys, subprocess, json, re, os
sys.path.insert(0, ".")
from tests import golden_io
for name, budget in (("accap_a3", 45000), ("trains15", 10000)):
    pb, info = golden_io.load(name)
    path = f"/tmp/{name}.tnf"
    golden_io.write_tnf(path, pb, info)
    r = subprocess.run(["turbo_b200/bin/turbo", "-s", "-t", str(budget), "-gpus", "$N", path], capture_output=True, text=True)
    st = dict(re.findall(r"^%%%mzn-stat: (\w+)=(.*)$", r.stdout, flags=re.M))
    print(json.dumps({"workload": name, "gpus": $N, "budget_ms": budget, "rc": r.returncode, "objective": st.get("objective"), "best_obj_time": st.get("best_obj_time"),
                      "nodes": st.get("nodes"), "solveTime": st.get("solveTime"), "fixpoint": st.get("fixpoint"), "stolen": st.get("eps_stolen_subproblems"),
                      "split": st.get("eps_split_subproblems"), "parts": st.get("eps_split_parts_solved"), "exhaustive": "==========" in r.stdout}), flush=True)
PY
