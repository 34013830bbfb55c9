#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -f -o gpurun_out/solve_active_final \
  python bench.py --fp wac1_active --steps 1 --warmup 0 --cutnodes 300 --no-cpu-baseline --no-fixpoint-leg > gpurun_out/ncu_active_final.log 2>&1
tail -2 gpurun_out/ncu_active_final.log
