#!/bin/bash
# Round 2, GPU visit 5: whole GPU suite and the default bench with the table in tensor memory, two CTAs per SM again;
# ncu: launch list, full capture of the solve kernel, DRAM traffic of a default-size launch.
mkdir -p gpurun_out
O=gpurun_out
( time timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 300 ) > $O/pytest_gpu_v5.log 2>&1; tail -6 $O/pytest_gpu_v5.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
B="--steps 5 --warmup 3 --no-cpu-baseline --strong-ms 0 --e2e-steps 5"
for t in 1 0; do
  TB_TMEM=$t timeout 300 python bench.py $B > $O/ab5_tmem${t}_trains15.json 2> $O/ab5_tmem${t}_trains15.err
  TB_TMEM=$t timeout 300 python bench.py $B --workload simplified:accap_a3 --no-fixpoint-leg > $O/ab5_tmem${t}_accap.json 2> $O/ab5_tmem${t}_accap.err
done
for f in $O/ab5_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    fk=d.get("fixpoint_kernel",{}); a=d.get("active_set",{})
    print(sys.argv[1].split("/")[-1], "blocks %d Gprop/s %.1f nodes/s %.0f frac %.3f fpshare %.2f e2e %.1f (%s) | fixpoint-alone %.1f | active nodes/s %.0f" % (d["config"]["num_blocks_per_gpu"], d["value"]/1e9, d["nodes_per_sec"], d["roofline"]["frac"], d["fixpoint_time_share"] or 0, d["e2e"]["value"]/1e9, {k: round(v,1) for k,v in d["e2e"]["split_ms_per_step"].items()}, fk.get("propagations_per_sec",0)/1e9, a.get("nodes_per_sec",0)))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
done
timeout 600 python bench.py > $O/bench_default_v5.json 2> $O/bench_default_v5.err; head -c 1800 $O/bench_default_v5.json; echo
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm_v5.json 2>> $O/bench_default_v5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_v5.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --strong-ms 0 --e2e-steps 1 > $O/ncu_launches_v5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -f -o $O/solve_v3tm_trains15 \
  python bench.py --steps 1 --warmup 0 --cutnodes 300 --no-cpu-baseline --no-fixpoint-leg --strong-ms 0 --e2e-steps 0 > $O/ncu_solve_tm.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,lts__t_bytes.sum --clock-control none -k regex:solve_kernel -c 1 --csv --log-file $O/traffic_default_launch.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-fixpoint-leg --strong-ms 0 --e2e-steps 0 > $O/ncu_traffic.log 2>&1
grep solve_kernel $O/traffic_default_launch.csv | cut -d, -f13- | head -8
