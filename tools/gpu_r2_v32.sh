#!/bin/bash
# Round 2, GPU visit 32 (1 GPU): last check of HEAD - whole GPU suite, smoke(), the default bench line.
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/v32_tests.txt 2>&1; tail -3 $O/v32_tests.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/v32_smoke.txt 2>&1; tail -1 $O/v32_smoke.txt
timeout 300 python bench.py --no-cpu-baseline > $O/v32_bench.json 2> $O/v32_bench.err; python -c "
import json; d=json.load(open('gpurun_out/v32_bench.json')); print(round(d['value']/1e9,1), round(d['nodes_per_sec']), round(d['e2e']['value']/1e9,1), round(d['roofline']['frac'],3), d['clocks'])"
