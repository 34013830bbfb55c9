#!/bin/bash
# Round 2, GPU visit 13 (2 GPUs): the tail-splitting pool shared between the GPUs - the multi-GPU tests, the driver with
# -gpus 2 on accap_a3 with and without the shared tail, and the absolute-address table walk on accap_a3 (TCN_SHARED).
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > $O/v13_multi_gpu_tests.txt 2>&1; tail -5 $O/v13_multi_gpu_tests.txt
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "dive_subproblems or single_block_trace or tail_splitting" > $O/v13_configs_tests.txt 2>&1; tail -3 $O/v13_configs_tests.txt
python - <<PY
import sys, subprocess, json, re, os
sys.path.insert(0, ".")
from tests import golden_io
pb, info = golden_io.load("accap_a3")
path = "/tmp/accap_a3.tnf"
golden_io.write_tnf(path, pb, info)
for g, share in ((1, "1"), ($N, "1"), ($N, "0")):
    env = dict(os.environ, TB_SHARE_SPLIT=share)
    r = subprocess.run(["turbo_b200/bin/turbo", "-s", "-t", "15000", "-gpus", str(g), path], capture_output=True, text=True, env=env)
    st = dict(re.findall(r"^%%%mzn-stat: (\w+)=(.*)$", r.stdout, flags=re.M))
    print(json.dumps({"workload": "accap_a3", "gpus": g, "share_split": share, "budget_ms": 15000, "rc": r.returncode, "objective": st.get("objective"),
                      "best_obj_time": st.get("best_obj_time"), "nodes": st.get("nodes"), "solveTime": st.get("solveTime"), "stolen": st.get("eps_stolen_subproblems"),
                      "split": st.get("eps_split_subproblems"), "parts": st.get("eps_split_parts_solved"), "exhaustive": "==========" in r.stdout}), flush=True)
PY
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --workload simplified:accap_a3 --no-cpu-baseline > $O/v13_bench_accap_n1.json 2> $O/v13_bench.err
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > $O/v13_bench_n1.json 2>> $O/v13_bench.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/v13_bench*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "Gprop/s %.1f nodes/s %.0f e2e %.1f frac %.3f fixpoint %s" % (d["value"] / 1e9, d["nodes_per_sec"], d["e2e"]["value"] / 1e9, d["roofline"]["frac"], d.get("fixpoint_kernel")))
    except Exception as e:
        print(f, "ERR", e)
PY
