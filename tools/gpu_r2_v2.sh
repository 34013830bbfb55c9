#!/bin/bash
# Round 2, GPU visit 2: A/B of the dense-fixpoint variants (prefetch pinned / two rows per lane), their parity,
# validation of the shared-memory peak microbenchmark (SM cycle counter + ncu), where tb_create / tb_destroy spend
# their time, the streaming tests.
mkdir -p gpurun_out
O=gpurun_out
V=$PWD/turbo_b200/variants
python -c "from turbo_b200 import engine; print(engine.measure_smem_peak(0))" > $O/smem_peak_v2.txt 2>&1; cat $O/smem_peak_v2.txt
( time timeout -k 10 600 python -m pytest tests/test_gpu_stream.py tests/test_gpu_configs.py -q --timeout 300 ) > $O/pytest_gpu_v2.log 2>&1; tail -6 $O/pytest_gpu_v2.log
for v in pin u2pin; do
  ( TURBO_B200_LIB=$V/libturbo_b200_$v.so timeout -k 10 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_active.py -q -x --timeout 300 ) > $O/pytest_variant_$v.log 2>&1
  echo "variant $v: $(tail -1 $O/pytest_variant_$v.log)"
done
B="--steps 5 --warmup 3 --no-cpu-baseline --strong-ms 0 --e2e-steps 3"
for v in default pin u2 u2pin v2; do
  if [ $v = default ]; then unset TURBO_B200_LIB; else export TURBO_B200_LIB=$V/libturbo_b200_$v.so; fi
  timeout 300 python bench.py $B > $O/ab2_${v}_trains15.json 2> $O/ab2_${v}_trains15.err
  timeout 300 python bench.py $B --workload simplified:accap_a3 --no-fixpoint-leg > $O/ab2_${v}_accap.json 2> $O/ab2_${v}_accap.err
  timeout 300 python bench.py $B --workload simplified:example_wordpress7_500 --no-fixpoint-leg > $O/ab2_${v}_wordpress.json 2> $O/ab2_${v}_wordpress.err
done
unset TURBO_B200_LIB
for f in $O/ab2_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    fk=d.get("fixpoint_kernel",{}); a=d.get("active_set",{})
    print(sys.argv[1].split("/")[-1], "Gprop/s %.1f nodes/s %.0f frac_nominal %.3f fpshare %.2f e2e %.1f | fixpoint-alone %.1f | active nodes/s %.0f" % (d["value"]/1e9, d["nodes_per_sec"], d["roofline"]["frac_of_nominal"], d["fixpoint_time_share"] or 0, d["e2e"]["value"]/1e9, fk.get("propagations_per_sec",0)/1e9, a.get("nodes_per_sec",0)))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
done
TB_TRACE_TIMING=1 python - > $O/timing_trace.txt 2>&1 <<'P'
from tests import golden_io
from turbo_b200 import engine
pb,_=golden_io.load_simplified_problem("trains15")
for i in range(3):
    print("--- round", i, flush=True)
    with engine.Solver(pb, cutnodes=200, subproblems_power=17) as s:
        s.solve()
P
cat $O/timing_trace.txt | tail -32
timeout 300 ncu --set full --clock-control none -k regex:smem_stream -c 1 -f -o $O/smem_stream python -c "from turbo_b200 import engine; print(engine.measure_smem_peak(0))" > $O/ncu_smem.log 2>&1
ncu -i $O/smem_stream.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for k in ('gpu__time_duration.sum','sm__cycles_elapsed.max','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','smsp__inst_executed_op_shared_ld.sum','launch__grid_size','smsp__cycles_active.avg'):
    if k in h: print(k, rows[2][h.index(k)], rows[1][h.index(k)])
"
