#!/bin/bash
# Round 2, GPU visit 31 (1 GPU): sweep-epilogue trim against the build before it (same box, alternating).
mkdir -p gpurun_out
O=gpurun_out
: > $O/v31_ab.jsonl
for rep in 1 2; do
for v in base new; do
  for w in trains15 accap_a3; do
    if [ $v = base ]; then export TURBO_B200_LIB=$PWD/turbo_b200/variants/libturbo_b200_base.so; else unset TURBO_B200_LIB; fi
    timeout 200 python bench.py --gpus 1 --steps 5 --warmup 3 --workload simplified:$w --sub 19 --no-cpu-baseline --strong-ms 0 --e2e-steps 3 > $O/v31_tmp.json 2> $O/v31_tmp.err
    python - $v $w <<'PY' | tee -a gpurun_out/v31_ab.jsonl
import json, sys
try:
    d = json.load(open("gpurun_out/v31_tmp.json")); k = d.get("fixpoint_kernel") or {}
    print(json.dumps({"build": sys.argv[1], "workload": sys.argv[2], "Gprop_s": round(d["value"] / 1e9, 1), "nodes_per_sec": round(d["nodes_per_sec"]), "fixpoint_alone_G": round(k.get("propagations_per_sec", 0) / 1e9, 1)}))
except Exception as e:
    print(json.dumps({"build": sys.argv[1], "workload": sys.argv[2], "error": str(e), "stderr": open("gpurun_out/v31_tmp.err").read()[-300:]}))
PY
  done
done
done
