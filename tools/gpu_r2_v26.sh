#!/bin/bash
# Round 2, GPU visit 26 (2 GPUs): the final build on two GPUs - multi-GPU tests, bench at N = 2, accap_a3 bench at N = 1.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_cli_gpu.py -m gpu -q > $O/v26_tests.txt 2>&1; tail -3 $O/v26_tests.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 5 --warmup 3 > $O/v26_bench_n2.json 2> $O/v26_bench_n2.err
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --workload simplified:accap_a3 --no-cpu-baseline --strong-ms 0 > $O/v26_bench_accap_n1.json 2> $O/v26_bench_accap.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/v26_bench*.json")):
    try:
        d = json.load(open(f)); s = d.get("strong_scaling") or {}; c = d["config"]
        print(f.split("/")[-1], "n_gpus", d["n_gpus"], c["num_blocks_per_gpu"], "x", c["threads_per_block"], "sub", c["subproblems_power"], "Gprop/s %.1f nodes/s %.0f e2e %.1f frac %.3f | strong nodes/s %.0f" % (d["value"] / 1e9, d["nodes_per_sec"], d["e2e"]["value"] / 1e9, d["roofline"]["frac"], s.get("nodes_per_sec", 0)))
    except Exception as e:
        print(f, "ERR", e)
PY
