#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q ) > $O/pytest_preload.log 2>&1
tail -4 $O/pytest_preload.log
B="python bench.py --no-cpu-baseline --no-fixpoint-leg"
run() { echo "== $*" >> $O/exp15.log; timeout 300 "$@" >> $O/exp15.log 2>> $O/exp15.err; }
for w in simplified:trains15 trains15 simplified:example_wordpress7_500 simplified:accap_a3; do
  run env TURBO_B200_LIB=$PWD/turbo_b200/variants/libturbo_b200_base.so $B --workload $w
  run $B --workload $w
done
echo "== fixpoint kernel trains15 base" >> $O/exp15.log
TURBO_B200_LIB=$PWD/turbo_b200/variants/libturbo_b200_base.so timeout 120 python tools/fixpoint_bench.py --workload trains15 >> $O/exp15.log 2>> $O/exp15.err
echo "== fixpoint kernel trains15 new" >> $O/exp15.log
timeout 120 python tools/fixpoint_bench.py --workload trains15 >> $O/exp15.log 2>> $O/exp15.err
python - <<'PY'
import json
for line in open("gpurun_out/exp15.log"):
    line = line.strip()
    if line.startswith("{"):
        d = json.loads(line)
        if "config" in d:
            c = d["config"]
            print("   %s tpb %d blocks %d | Gprop/s %.1f nodes/s %.0f frac %.4f" % (c["memory_configuration"], c["threads_per_block"], c["num_blocks_per_gpu"], d["value"] / 1e9, d["nodes_per_sec"], d["roofline"]["frac"]))
        else:
            print("   fixpoint kernel Gprop/s %.1f frac %.4f ms %.3f" % (d["propagations_per_sec"] / 1e9, d["smem_frac"], d["kernel_ms"]))
    else:
        print(line)
PY
tail -3 $O/exp15.err
