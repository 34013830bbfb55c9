#!/bin/bash
# Round 2, GPU visit 14 (1 GPU): ncu of the solve kernel on accap_a3 (the small-network node rate).
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -f -o $O/solve_accap_v14 \
  python bench.py --steps 1 --warmup 0 --cutnodes 300 --workload simplified:accap_a3 --no-cpu-baseline --no-fixpoint-leg --strong-ms 0 --e2e-steps 0 > $O/ncu_solve_accap_v14.log 2>&1
tail -3 $O/ncu_solve_accap_v14.log
ls -la $O/solve_accap_v14.ncu-rep
