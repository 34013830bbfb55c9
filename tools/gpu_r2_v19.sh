#!/bin/bash
# Round 2, GPU visit 19 (1 GPU): block shapes, second pass: single-warp blocks at 16 / 20 / 24 per SM on accap_a3 (tensor
# memory holds part of the table only up to 16 per SM), thread counts on trains15.
mkdir -p gpurun_out
O=gpurun_out
: > $O/v19_shapes.jsonl
run() {
  timeout 200 python bench.py --gpus 1 --steps 3 --warmup 3 --workload $1 --mem $2 --tpb $3 --blocks $4 --no-cpu-baseline --strong-ms 0 --no-fixpoint-leg --e2e-steps 3 > $O/v19_tmp.json 2> $O/v19_tmp.err
  python - "$@" <<'PY' | tee -a gpurun_out/v19_shapes.jsonl
import json, sys
try:
    d = json.load(open("gpurun_out/v19_tmp.json")); c = d["config"]
    print(json.dumps({"workload": sys.argv[1], "mem_arg": sys.argv[2], "tpb_arg": sys.argv[3], "blocks_arg": sys.argv[4], "env": sys.argv[5] if len(sys.argv) > 5 else "", "memory_configuration": c["memory_configuration"], "threads_per_block": c["threads_per_block"], "blocks": c["num_blocks_per_gpu"],
                      "Gprop_s": round(d["value"] / 1e9, 1), "nodes_per_sec": round(d["nodes_per_sec"]), "fixpoint_time_share": round(d["fixpoint_time_share"], 3)}))
except Exception as e:
    print(json.dumps({"args": sys.argv[1:], "error": str(e), "stderr": open("gpurun_out/v19_tmp.err").read()[-300:]}))
PY
}
run simplified:accap_a3 store_shared 32 2368
TB_TMEM=0 run simplified:accap_a3 store_shared 32 2368 TB_TMEM=0
run simplified:accap_a3 store_shared 32 2960
run simplified:accap_a3 store_shared 32 3552
run simplified:accap_a3 store_shared 64 1776
run simplified:trains15 store_shared 256 0
run simplified:trains15 store_shared 512 0
run simplified:trains15 store_shared 1024 0
