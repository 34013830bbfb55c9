#!/bin/bash
# GPU visit 2: simplifier + CLI tests, benches of the simplified networks, placement / variant experiments.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_simplifier.py tests/test_cli_gpu.py -m gpu -x -q ) > $O/pytest_simp.log 2>&1
tail -3 $O/pytest_simp.log
timeout 600 python bench.py --workload simplified:trains15 > $O/bench_strains.json 2> $O/bench_strains.err
timeout 300 python bench.py --workload simplified:example_wordpress7_500 --no-cpu-baseline > $O/bench_swordpress.json 2> $O/bench_swordpress.err
timeout 300 python bench.py --workload simplified:accap_a3 --no-cpu-baseline > $O/bench_saccap.json 2> $O/bench_saccap.err
for w in simplified:trains15 simplified:accap_a3 simplified:example_wordpress7_500; do
  timeout 60 python tools/time_to_optimum.py $w --timeout-ms 20000 >> $O/tto_simplified.jsonl 2>> $O/tto.err
done
cat $O/tto_simplified.jsonl
for mem in global store_cluster; do
  echo "synthetic $mem" >> $O/exp.log
  timeout 120 python tools/fixpoint_bench.py --workload synthetic --repeat 2 --rounds 1 --mem $mem >> $O/exp.log 2>&1
done
for w in trains15 simplified:trains15; do
  echo "U1 $w" >> $O/exp.log
  timeout 120 python tools/fixpoint_bench.py --workload $w >> $O/exp.log 2>&1
  echo "U2 $w" >> $O/exp.log
  TURBO_B200_LIB=$PWD/turbo_b200/variants/libturbo_b200_u2.so timeout 120 python tools/fixpoint_bench.py --workload $w >> $O/exp.log 2>&1
  echo "U1 $w tpb512" >> $O/exp.log
  timeout 120 python tools/fixpoint_bench.py --workload $w --tpb 512 >> $O/exp.log 2>&1
done
cat $O/exp.log
