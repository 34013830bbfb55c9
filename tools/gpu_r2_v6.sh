#!/bin/bash
# Round 2, GPU visit 6 (2 GPUs): the multi-GPU tests (epochs, final gather, stealing, satisfaction stop), tail splitting,
# bench at N = 2 next to N = 1 on the same box (weak step, strong-scaling leg, NCCL gather), the driver with -gpus,
# accap_a3 towards a proof.
N=2
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/v6_gpus.txt 2>&1
( time timeout -k 10 900 python -m pytest tests/test_multi_gpu.py tests/test_gpu_stream.py "tests/test_gpu_configs.py::test_tail_splitting_keeps_status_and_optimum" tests/test_cli_gpu.py -m gpu -q --timeout 300 ) > $O/pytest_v6.log 2>&1
tail -12 $O/pytest_v6.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 5 --warmup 3 > $O/v6_bench_n$N.json 2> $O/v6_bench_n$N.err
tail -3 $O/v6_bench_n$N.err
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-fixpoint-leg > $O/v6_bench_n1.json 2>> $O/v6_bench_n$N.err
for w in simplified:accap_a3; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus $N --steps 5 --warmup 3 --workload $w --strong-sub 22 --no-cpu-baseline > $O/v6_bench_accap_n$N.json 2>> $O/v6_bench_n$N.err
  timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --workload $w --strong-sub 22 --no-cpu-baseline --no-fixpoint-leg > $O/v6_bench_accap_n1.json 2>> $O/v6_bench_n$N.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/v6_bench*.json")):
    try:
        d = json.load(open(f)); s = d.get("strong_scaling") or {}
        print(f.split("/")[-1], "n_gpus", d["n_gpus"], "Gprop/s %.1f nodes/s %.0f e2e %.1f best %s | strong: nodes/s %.0f solved %s stolen %s best %s t_best %.0f ms" % (
            d["value"] / 1e9, d["nodes_per_sec"], d["e2e"]["value"] / 1e9, d["best_objective"], s.get("nodes_per_sec", 0), s.get("subproblems_solved"), s.get("subproblems_stolen"), s.get("best_objective"), s.get("time_to_best_ms", 0)))
    except Exception as e:
        print(f, "ERR", e)
PY
# the driver: one process, -gpus 1 / 2, 10 s budget; then accap_a3 towards a proof on one GPU (tail splitting)
python - <<PY
import sys, subprocess, json, re, os
sys.path.insert(0, ".")
from tests import golden_io
for name in ("accap_a3", "trains15"):
    pb, info = golden_io.load(name)
    path = f"/tmp/{name}.tnf"
    golden_io.write_tnf(path, pb, info)
    for g in (1, $N):
        r = subprocess.run(["turbo_b200/bin/turbo", "-s", "-t", "10000", "-gpus", str(g), path], capture_output=True, text=True)
        st = dict(re.findall(r"^%%%mzn-stat: (\w+)=(.*)$", r.stdout, flags=re.M))
        print(json.dumps({"workload": name, "gpus": g, "rc": r.returncode, "objective": st.get("objective"), "best_obj_time": st.get("best_obj_time"),
                          "nodes": st.get("nodes"), "solveTime": st.get("solveTime"), "fixpoint": st.get("fixpoint"), "stolen": st.get("eps_stolen_subproblems"),
                          "split": st.get("eps_split_subproblems"), "exhaustive": "==========" in r.stdout}))
PY
for sub in -1 22; do
  timeout 200 python tools/time_to_optimum.py simplified:accap_a3 --timeout-ms 90000 --fp wac1 --sub $sub >> $O/v6_tto_accap.jsonl 2>> $O/v6_tto.err
done
cat $O/v6_tto_accap.jsonl | cut -c1-600
