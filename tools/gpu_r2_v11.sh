#!/bin/bash
# Round 2, GPU visit 11: absolute shared addresses in the tensor-memory words (33 instructions per visit): parity, bench.
mkdir -p gpurun_out
O=gpurun_out
( time timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 300 ) > $O/pytest_gpu_v11.log 2>&1; tail -6 $O/pytest_gpu_v11.log
B="--steps 5 --warmup 3 --no-cpu-baseline --strong-ms 0 --e2e-steps 3"
timeout 300 python bench.py $B > $O/ab11_abs_trains15.json 2> $O/ab11.err
timeout 300 python bench.py $B --workload trains15 --no-fixpoint-leg > $O/ab11_abs_trains15full.json 2>> $O/ab11.err
timeout 300 python bench.py $B --workload simplified:example_wordpress7_500 --no-fixpoint-leg > $O/ab11_mixed_wordpress.json 2>> $O/ab11.err
timeout 300 python bench.py $B --workload simplified:accap_a3 --no-fixpoint-leg > $O/ab11_tcn_accap.json 2>> $O/ab11.err
for f in $O/ab11_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    fk=d.get("fixpoint_kernel",{}); a=d.get("active_set",{})
    print(sys.argv[1].split("/")[-1], "blocks %d Gprop/s %.1f nodes/s %.0f frac %.3f fpshare %.2f e2e %.1f | fixpoint-alone %.1f (%.3f) | active nodes/s %.0f" % (d["config"]["num_blocks_per_gpu"], d["value"]/1e9, d["nodes_per_sec"], d["roofline"]["frac"], d["fixpoint_time_share"] or 0, d["e2e"]["value"]/1e9, fk.get("propagations_per_sec",0)/1e9, fk.get("smem_frac",0), a.get("nodes_per_sec",0)))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
done
