#!/bin/bash
# Round 2, GPU visit 12 (8 GPUs): bench at N = 8 next to N = 1 on the same box (weak step + fixed-pool strong-scaling
# leg + NCCL gather), accap_a3 fixed pool at N = 8 and N = 1, the driver with -gpus 8 (accap_a3 towards a proof, trains15).
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/v12_gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 5 --warmup 3 > $O/v12_bench_n$N.json 2> $O/v12_bench_n$N.err
tail -2 $O/v12_bench_n$N.err
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-fixpoint-leg > $O/v12_bench_n1.json 2>> $O/v12_bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus $N --steps 5 --warmup 3 --workload simplified:accap_a3 --strong-sub 24 --strong-ms 4000 --no-cpu-baseline > $O/v12_bench_accap_n$N.json 2>> $O/v12_bench_n$N.err
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --workload simplified:accap_a3 --strong-sub 24 --strong-ms 4000 --no-cpu-baseline --no-fixpoint-leg > $O/v12_bench_accap_n1.json 2>> $O/v12_bench_n$N.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/v12_bench*.json")):
    try:
        d = json.load(open(f)); s = d.get("strong_scaling") or {}
        print(f.split("/")[-1], "n_gpus", d["n_gpus"], "Gprop/s %.1f nodes/s %.0f e2e %.1f (%s) best %s | strong: nodes/s %.0f solved %s stolen %s best %s t_best %.0f ms" % (
            d["value"] / 1e9, d["nodes_per_sec"], d["e2e"]["value"] / 1e9, {k: round(v, 1) for k, v in d["e2e"]["split_ms_per_step"].items()}, d["best_objective"], s.get("nodes_per_sec", 0), s.get("subproblems_solved"), s.get("subproblems_stolen"), s.get("best_objective"), s.get("time_to_best_ms", 0)))
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
    for g in ($N,):
        r = subprocess.run(["turbo_b200/bin/turbo", "-s", "-t", str(budget), "-gpus", str(g), path], capture_output=True, text=True)
        st = dict(re.findall(r"^%%%mzn-stat: (\w+)=(.*)$", r.stdout, flags=re.M))
        print(json.dumps({"workload": name, "gpus": g, "budget_ms": budget, "rc": r.returncode, "objective": st.get("objective"), "best_obj_time": st.get("best_obj_time"),
                          "nodes": st.get("nodes"), "solveTime": st.get("solveTime"), "fixpoint": st.get("fixpoint"), "stolen": st.get("eps_stolen_subproblems"),
                          "split": st.get("eps_split_subproblems"), "parts": st.get("eps_split_parts_solved"), "exhaustive": "==========" in r.stdout}), flush=True)
PY
