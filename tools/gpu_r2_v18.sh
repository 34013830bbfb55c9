#!/bin/bash
# Round 2, GPU visit 18 (1 GPU): block shapes for the small networks (accap_a3): threads per block x placement.
mkdir -p gpurun_out
O=gpurun_out
: > $O/v18_shapes.jsonl
for cfg in "auto 0" "tcn_shared 64" "tcn_shared 32" "tcn_shared 256" "store_shared 128" "store_shared 64" "store_shared 32"; do
  set -- $cfg
  timeout 200 python bench.py --gpus 1 --steps 3 --warmup 3 --workload simplified:accap_a3 --mem $1 --tpb $2 --no-cpu-baseline --strong-ms 0 --no-fixpoint-leg --e2e-steps 3 > $O/v18_tmp.json 2> $O/v18_tmp.err
  python - "$1" "$2" <<'PY' | tee -a gpurun_out/v18_shapes.jsonl
import json, sys
try:
    d = json.load(open("gpurun_out/v18_tmp.json")); c = d["config"]
    print(json.dumps({"mem_arg": sys.argv[1], "tpb_arg": sys.argv[2], "memory_configuration": c["memory_configuration"], "threads_per_block": c["threads_per_block"], "blocks": c["num_blocks_per_gpu"],
                      "Gprop_s": round(d["value"] / 1e9, 1), "nodes_per_sec": round(d["nodes_per_sec"]), "fixpoint_time_share": round(d["fixpoint_time_share"], 3)}))
except Exception as e:
    print(json.dumps({"mem_arg": sys.argv[1], "tpb_arg": sys.argv[2], "error": str(e), "stderr": open("gpurun_out/v18_tmp.err").read()[-300:]}))
PY
done
