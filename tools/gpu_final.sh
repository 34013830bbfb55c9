#!/bin/bash
# Final GPU visit of a round: the whole GPU suite, smoke(), the default bench line and the reference arm, the ncu launch
# list of the same command. (ncu --set full captures and time-to-optimum runs: see profiles/README.md for the commands.)
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/pytest_gpu_final.log 2>&1
tail -4 $O/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > $O/bench_final.json 2> $O/bench_final.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2>> $O/bench_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_final.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1
head -c 1200 $O/bench_final.json
