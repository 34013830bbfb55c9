#!/bin/bash
# Round 2, GPU visit 3: the propagator table in tensor memory (parity + A/B), the three shared-memory peak modes,
# e2e after the create/destroy fixes, ncu captures (with TMEM; default-size launch for the DRAM traffic), the
# BASELINE.md §3 report table with a 10 s budget.
mkdir -p gpurun_out
O=gpurun_out
for m in 2 1 0; do TB_SMEM_PEAK_MODE=$m python -c "from turbo_b200 import engine; print('mode $m', engine.measure_smem_peak(0))"; done > $O/smem_peak_v3.txt 2>&1; cat $O/smem_peak_v3.txt
( time timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_stream.py tests/test_known_answers.py -q --timeout 300 -m gpu ) > $O/pytest_gpu_v3.log 2>&1; tail -8 $O/pytest_gpu_v3.log
B="--steps 5 --warmup 3 --no-cpu-baseline --strong-ms 0 --e2e-steps 5"
for t in 1 0; do
  TB_TMEM=$t timeout 300 python bench.py $B > $O/ab3_tmem${t}_trains15.json 2> $O/ab3_tmem${t}_trains15.err
  TB_TMEM=$t timeout 300 python bench.py $B --workload trains15 --no-fixpoint-leg > $O/ab3_tmem${t}_trains15full.json 2> $O/ab3_tmem${t}_trains15full.err
  TB_TMEM=$t timeout 300 python bench.py $B --workload simplified:example_wordpress7_500 --no-fixpoint-leg > $O/ab3_tmem${t}_wordpress.json 2> $O/ab3_tmem${t}_wordpress.err
done
for f in $O/ab3_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    fk=d.get("fixpoint_kernel",{}); a=d.get("active_set",{})
    print(sys.argv[1].split("/")[-1], "Gprop/s %.1f nodes/s %.0f frac_nominal %.3f fpshare %.2f e2e %.1f (%s) | fixpoint-alone %.1f | active nodes/s %.0f" % (d["value"]/1e9, d["nodes_per_sec"], d["roofline"]["frac_of_nominal"], d["fixpoint_time_share"] or 0, d["e2e"]["value"]/1e9, {k: round(v,1) for k,v in d["e2e"]["split_ms_per_step"].items()}, fk.get("propagations_per_sec",0)/1e9, a.get("nodes_per_sec",0)))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -f -o $O/solve_v3tm_trains15 \
  python bench.py --steps 1 --warmup 0 --cutnodes 300 --no-cpu-baseline --no-fixpoint-leg --strong-ms 0 --e2e-steps 0 > $O/ncu_solve_tm.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,lts__t_bytes.sum --clock-control none -k regex:solve_kernel -c 1 --csv --log-file $O/traffic_default_launch.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-fixpoint-leg --strong-ms 0 --e2e-steps 0 > $O/ncu_traffic.log 2>&1
grep solve_kernel $O/traffic_default_launch.csv | cut -d, -f5,13- | head -8
timeout 300 ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max,smsp__inst_executed_op_shared_ld.sum --clock-control none -k regex:smem_stream -c 1 --csv --log-file $O/smem_gather_ncu.csv python -c "from turbo_b200 import engine; print(engine.measure_smem_peak(0))" > /dev/null 2>&1
grep smem_stream $O/smem_gather_ncu.csv | cut -d, -f13- | head
timeout 900 python bench.py --report 1,2,3 --report-ms 10000 > $O/report_n1_10s.txt 2> $O/report_n1.err
tail -16 $O/report_n1_10s.txt
