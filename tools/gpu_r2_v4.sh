#!/bin/bash
# Round 2, GPU visit 4 (short): why the occupancy API reports one CTA per SM for the kernels that carry tensor-memory
# code, what the hardware really does with two, and the corrected shared-memory peak benchmark.
mkdir -p gpurun_out
O=gpurun_out
V=$PWD/turbo_b200/variants
for m in 2 1 0; do TB_SMEM_PEAK_MODE=$m python -c "from turbo_b200 import engine; print('mode $m', engine.measure_smem_peak(0))"; done > $O/smem_peak_v4.txt 2>&1; cat $O/smem_peak_v4.txt
cat > /tmp/occ.py <<'P'
from tests import golden_io
from turbo_b200 import engine
pb,_=golden_io.load_simplified_problem("trains15")
with engine.Solver(pb, subproblems_power=17) as s:
    print(s.config()["num_blocks"], s.config()["blocks_per_sm"])
P
TB_TRACE_TIMING=1 python /tmp/occ.py 2>&1 | grep -E "tb config|^[0-9]"
TB_TRACE_TIMING=1 TURBO_B200_LIB=$V/libturbo_b200_notm.so python /tmp/occ.py 2>&1 | grep -E "tb config|^[0-9]"
B="--steps 5 --warmup 3 --no-cpu-baseline --strong-ms 0 --e2e-steps 3"
TB_IGNORE_OCCUPANCY_API=1 timeout 300 python bench.py $B > $O/ab4_tmem1_ignore_trains15.json 2> $O/ab4_a.err
TB_IGNORE_OCCUPANCY_API=1 TB_TMEM=0 timeout 300 python bench.py $B --no-fixpoint-leg > $O/ab4_tmem0_ignore_trains15.json 2> $O/ab4_b.err
TURBO_B200_LIB=$V/libturbo_b200_notm.so timeout 300 python bench.py $B --no-fixpoint-leg > $O/ab4_notm_trains15.json 2> $O/ab4_c.err
for f in $O/ab4_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    fk=d.get("fixpoint_kernel",{})
    print(sys.argv[1].split("/")[-1], "blocks %d Gprop/s %.1f nodes/s %.0f frac %.3f frac_nominal %.3f fpshare %.2f e2e %.1f | fixpoint-alone %.1f" % (d["config"]["num_blocks_per_gpu"], d["value"]/1e9, d["nodes_per_sec"], d["roofline"]["frac"], d["roofline"]["frac_of_nominal"], d["fixpoint_time_share"] or 0, d["e2e"]["value"]/1e9, fk.get("propagations_per_sec",0)/1e9))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
done
tail -3 $O/ab4_a.err
