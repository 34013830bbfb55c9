#!/bin/bash
# Round 2, GPU visit 1: GPU suite with the third-generation dense fixpoint, A/B of the kernel variants, the other
# configurations, ncu captures (solve kernel on trains15, cluster kernel on the synthetic network), long runs towards
# the optima of the headline instances.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
( time timeout -k 10 1200 python -m pytest tests -m gpu -q --timeout 300 ) > $O/pytest_gpu_v1.log 2>&1
tail -15 $O/pytest_gpu_v1.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -c "from turbo_b200 import engine; print(engine.device_info(0)); print(engine.measure_smem_peak(0))" > $O/smem_peak.txt 2>&1
cat $O/smem_peak.txt
B="--steps 5 --warmup 3 --no-cpu-baseline --strong-ms 0 --e2e-steps 3"
for v in default v2 u2; do
  if [ $v = default ]; then unset TURBO_B200_LIB; else export TURBO_B200_LIB=$PWD/turbo_b200/variants/libturbo_b200_$v.so; fi
  timeout 300 python bench.py $B > $O/ab_${v}_trains15.json 2> $O/ab_${v}_trains15.err
  timeout 300 python bench.py $B --workload simplified:accap_a3 --no-fixpoint-leg > $O/ab_${v}_accap.json 2> $O/ab_${v}_accap.err
  timeout 300 python bench.py $B --workload simplified:example_wordpress7_500 --no-fixpoint-leg > $O/ab_${v}_wordpress.json 2> $O/ab_${v}_wordpress.err
done
unset TURBO_B200_LIB
for f in $O/ab_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    fk=d.get("fixpoint_kernel",{}); a=d.get("active_set",{})
    print(sys.argv[1].split("/")[-1], "Gprop/s %.1f nodes/s %.0f frac %.3f fpshare %.2f e2e %.1f | fixpoint-alone %.1f | active nodes/s %.0f" % (d["value"]/1e9, d["nodes_per_sec"], d["roofline"]["frac"], d["fixpoint_time_share"] or 0, d["e2e"]["value"]/1e9, fk.get("propagations_per_sec",0)/1e9, a.get("nodes_per_sec",0)))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
done
TB_L2_PERSIST=0 timeout 300 python bench.py $B --no-fixpoint-leg > $O/ab_default_nopersist_trains15.json 2>> $O/ab.err
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
head -c 1500 $O/bench_default.json; echo
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2>> $O/bench_default.err
for m in store_cluster global; do
  timeout 300 python tools/fixpoint_bench.py --workload synthetic:100000:1000000 --mem $m --repeat 5 --rounds 2 >> $O/synthetic_fixpoint.jsonl 2>> $O/synthetic.err
done
cat $O/synthetic_fixpoint.jsonl | cut -c1-600
timeout 300 python bench.py $B --workload example_wordpress7_500 --no-fixpoint-leg > $O/bench_wordpress_unsimplified.json 2>> $O/ab.err
# ncu: launch list of the default command, full captures
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_v1.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --strong-ms 0 --e2e-steps 1 > $O/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -f -o $O/solve_v3_trains15 \
  python bench.py --steps 1 --warmup 0 --cutnodes 300 --no-cpu-baseline --no-fixpoint-leg --strong-ms 0 --e2e-steps 0 > $O/ncu_solve.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:propagate_kernel -c 1 -f -o $O/propagate_cluster_synthetic \
  python tools/fixpoint_bench.py --workload synthetic:100000:1000000 --mem store_cluster --repeat 2 --rounds 0 > $O/ncu_cluster.log 2>&1
# long runs towards the optima
for w in simplified:accap_a3 simplified:trains15 simplified:example_wordpress7_500; do
  timeout 100 python tools/time_to_optimum.py $w --timeout-ms 45000 --fp wac1_active >> $O/tto_v1.jsonl 2>> $O/tto.err
done
timeout 100 python tools/time_to_optimum.py simplified:accap_a3 --timeout-ms 45000 --fp wac1 --sub 24 >> $O/tto_v1.jsonl 2>> $O/tto.err
cat $O/tto_v1.jsonl | cut -c1-500
