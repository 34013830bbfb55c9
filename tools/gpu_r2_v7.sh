#!/bin/bash
# Round 2, GPU visit 7: cluster locality (parity, synthetic and unsimplified wordpress A/B, ncu), tail-splitting tests,
# whole GPU suite.
mkdir -p gpurun_out
O=gpurun_out
( time timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 300 ) > $O/pytest_gpu_v7.log 2>&1; tail -8 $O/pytest_gpu_v7.log
for loc in 1 0; do
  for m in store_cluster; do
    TB_CLUSTER_LOCALITY=$loc TB_TRACE_TIMING=1 timeout 300 python tools/fixpoint_bench.py --workload synthetic:100000:1000000 --mem $m --repeat 5 --rounds 2 >> $O/v7_synthetic_loc$loc.jsonl 2> $O/v7_synth_loc$loc.err
    grep "tb config" $O/v7_synth_loc$loc.err | tail -1
  done
  TB_CLUSTER_LOCALITY=$loc timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --strong-ms 0 --e2e-steps 2 --workload example_wordpress7_500 --no-fixpoint-leg > $O/v7_wordpress_unsimplified_loc$loc.json 2>> $O/v7.err
  TB_CLUSTER_LOCALITY=$loc timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --strong-ms 0 --e2e-steps 2 --workload synthetic:100000:1000000 --no-fixpoint-leg > $O/v7_synthetic_bench_loc$loc.json 2>> $O/v7.err
done
cat $O/v7_synthetic_loc*.jsonl | cut -c1-420
for f in $O/v7_wordpress_unsimplified_loc*.json $O/v7_synthetic_bench_loc*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1].split("/")[-1], d["config"]["memory_configuration"], "Gprop/s %.1f nodes/s %.0f frac %.3f e2e %.1f" % (d["value"]/1e9, d["nodes_per_sec"], d["roofline"]["frac"], d["e2e"]["value"]/1e9))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:propagate_kernel -c 1 -f -o $O/propagate_cluster_synthetic_loc \
  python tools/fixpoint_bench.py --workload synthetic:100000:1000000 --mem store_cluster --repeat 2 --rounds 0 > $O/ncu_cluster_loc.log 2>&1
