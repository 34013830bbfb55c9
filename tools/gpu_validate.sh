#!/bin/bash
# One GPU box visit: GPU test suite, the default bench line, the other BASELINE configs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 300 python bench.py --workload accap_a3 --no-cpu-baseline > gpurun_out/bench_accap.json 2> gpurun_out/bench_accap.err
timeout 300 python bench.py --workload example_wordpress7_500 --no-cpu-baseline > gpurun_out/bench_wordpress.json 2> gpurun_out/bench_wordpress.err
timeout 300 python bench.py --workload synthetic --steps 2 --warmup 1 --cutnodes 50 --no-cpu-baseline > gpurun_out/bench_synth.json 2> gpurun_out/bench_synth.err
tail -3 gpurun_out/pytest_gpu.log
head -c 600 gpurun_out/bench_default.json
for w in trains15 accap_a3 example_wordpress7_500; do
  timeout 60 python tools/time_to_optimum.py $w --timeout-ms 20000 >> gpurun_out/tto.jsonl 2>> gpurun_out/tto.err
done
cat gpurun_out/tto.jsonl
