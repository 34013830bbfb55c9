#!/bin/bash
# Final GPU visit of the round: the whole GPU suite, the default bench line, the ncu launch list of the same command,
# one full ncu capture of the solve kernel, time-to-optimum runs.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu_final.log 2>&1
tail -4 $O/pytest_gpu_final.log
timeout 600 python bench.py > $O/bench_final.json 2> $O/bench_final.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2>> $O/bench_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_final.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -f -o $O/solve_dense_snap_strains \
  python bench.py --steps 1 --warmup 0 --cutnodes 300 --no-cpu-baseline --no-fixpoint-leg > $O/ncu_solve_final.log 2>&1
for w in simplified:trains15 simplified:example_wordpress7_500 simplified:accap_a3; do
  for fp in wac1 wac1_active; do
    timeout 60 python tools/time_to_optimum.py $w --timeout-ms 10000 --fp $fp >> $O/tto_final.jsonl 2>> $O/tto.err
  done
done
cat $O/tto_final.jsonl | cut -c1-420
head -c 1500 $O/bench_final.json
