#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python tools/sweep.py --timeout-ms 5000 --out $O/sweep_wac1.md > $O/sweep_wac1.log 2>&1
tail -3 $O/sweep_wac1.log
timeout 900 python tools/sweep.py --timeout-ms 5000 --fp wac1_active --out $O/sweep_wac1_active.md > $O/sweep_wac1_active.log 2>&1
tail -3 $O/sweep_wac1_active.log
cat $O/sweep_wac1.md | head -45
