#!/bin/bash
# Round 2, GPU visit 15 (1 GPU): contiguous chunk ranges per warp on the TCN_SHARED tier - parity, then accap_a3.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q > $O/v15_tests.txt 2>&1; tail -3 $O/v15_tests.txt
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --workload simplified:accap_a3 --no-cpu-baseline --strong-ms 0 > $O/v15_bench_accap_n1.json 2> $O/v15_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/v15_bench_accap_n1.json"))
print("accap Gprop/s %.1f nodes/s %.0f frac %.3f fp_share %.3f fixpoint-alone %.1f G" % (d["value"] / 1e9, d["nodes_per_sec"], d["roofline"]["frac"], d["fixpoint_time_share"], d["fixpoint_kernel"]["propagations_per_sec"] / 1e9))
PY
