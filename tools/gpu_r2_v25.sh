#!/bin/bash
# Round 2, GPU visit 25 (1 GPU): is a partly-in-tensor-memory table worth it for few-warp blocks? accap_a3, 64 and 128
# threads per block at 16 / 8 blocks per SM, TB_TMEM on and off; pat13 (unsimplified) through the same bench.
mkdir -p gpurun_out
O=gpurun_out
: > $O/v25_tmem_mixed.jsonl
run() {
  timeout 200 python bench.py --gpus 1 --steps 3 --warmup 3 --workload $1 --mem $2 --tpb $3 --blocks $4 --no-cpu-baseline --strong-ms 0 --no-fixpoint-leg --e2e-steps 3 > $O/v25_tmp.json 2> $O/v25_tmp.err
  python - "$@" "TB_TMEM=${TB_TMEM:-1}" <<'PY' | tee -a gpurun_out/v25_tmem_mixed.jsonl
import json, sys
try:
    d = json.load(open("gpurun_out/v25_tmp.json")); c = d["config"]
    print(json.dumps({"workload": sys.argv[1], "mem_arg": sys.argv[2], "tpb_arg": sys.argv[3], "blocks_arg": sys.argv[4], "env": sys.argv[5], "memory_configuration": c["memory_configuration"], "threads_per_block": c["threads_per_block"], "blocks": c["num_blocks_per_gpu"],
                      "Gprop_s": round(d["value"] / 1e9, 1), "nodes_per_sec": round(d["nodes_per_sec"]), "fixpoint_time_share": round(d["fixpoint_time_share"], 3)}))
except Exception as e:
    print(json.dumps({"args": sys.argv[1:], "error": str(e), "stderr": open("gpurun_out/v25_tmp.err").read()[-300:]}))
PY
}
run simplified:accap_a3 store_shared 64 2368
TB_TMEM=0 run simplified:accap_a3 store_shared 64 2368
run simplified:accap_a3 store_shared 128 1184
TB_TMEM=0 run simplified:accap_a3 store_shared 128 1184
run pat13 auto 0 0
TB_TMEM=0 run pat13 auto 0 0
run pat9 auto 0 0
TB_TMEM=0 run pat9 auto 0 0
