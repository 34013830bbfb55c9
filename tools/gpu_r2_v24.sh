#!/bin/bash
# Round 2, GPU visit 24 (1 GPU): the final build - whole GPU suite, smoke(), the default bench and the reference arm, the
# ncu launch list of the bench command.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/v24_tests.txt 2>&1; tail -3 $O/v24_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/v24_smoke.txt 2>&1; tail -1 $O/v24_smoke.txt
timeout 600 python bench.py > $O/v24_bench.json 2> $O/v24_bench.err; tail -c 600 $O/v24_bench.json
timeout 600 python bench.py --impl reference > $O/v24_bench_reference.json 2> $O/v24_bench_reference.err; tail -c 400 $O/v24_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_v24.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --strong-ms 0 --e2e-steps 1 > $O/ncu_launches_v24.log 2>&1
wc -l $O/launches_v24.csv
