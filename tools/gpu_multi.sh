#!/bin/bash
# usage: tools/gpu_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/multi_gpus_$N.txt 2>&1
if [ "$N" = "2" ]; then
  ( time timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > $O/pytest_multi_$N.log 2>&1
  tail -4 $O/pytest_multi_$N.log
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err
tail -3 $O/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus $N --steps 5 --warmup 3 --workload simplified:accap_a3 > $O/bench_accap_n$N.json 2>> $O/bench_n$N.err
timeout 120 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-fixpoint-leg > $O/bench_n1_samebox_$N.json 2>> $O/bench_n$N.err
# time-to-optimum with bound sharing across the N GPUs through the driver (one process, tb_link_peers)
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
                          "nodes": st.get("nodes"), "solveTime": st.get("solveTime"), "exhaustive": "==========" in r.stdout}))
PY
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench*_n*.json")):
    try:
        d = json.load(open(f))
        print(f, "n_gpus", d["n_gpus"], "Gprop/s %.1f nodes/s %.0f ms/step %.1f frac %.4f" % (d["value"] / 1e9, d["nodes_per_sec"], d["ms_per_step"], d["roofline"]["frac"]))
    except Exception as e:
        print(f, "ERR", e)
PY
