#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/sweep.py --timeout-ms 5000 --out gpurun_out/sweep_wac1.md > gpurun_out/sweep_wac1.log 2>&1
tail -2 gpurun_out/sweep_wac1.log | cut -c1-200
timeout 600 python tools/sweep.py --timeout-ms 5000 --fp wac1_active --out gpurun_out/sweep_wac1_active.md > gpurun_out/sweep_wac1_active.log 2>&1
tail -2 gpurun_out/sweep_wac1_active.log | cut -c1-200
