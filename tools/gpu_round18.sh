#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_active.py tests/test_multi_gpu.py -m gpu -x -q ) > $O/pytest_r18.log 2>&1
tail -4 $O/pytest_r18.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
  bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_final_n2.json 2> $O/bench_final_n2.err
echo "stdout lines: $(wc -l < $O/bench_final_n2.json)"; head -c 300 $O/bench_final_n2.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 \
  bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > $O/bench_ref_n2.json 2>> $O/bench_final_n2.err
echo "ref stdout lines: $(wc -l < $O/bench_ref_n2.json)"; head -c 200 $O/bench_ref_n2.json; echo
turbo_b200/bin/turbo -s -v tests/data/tiny.fzn | grep -E "start-up|solveTime|initTime" 
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_final_n2.json") if l.startswith("{")][-1])
print("n_gpus", d["n_gpus"], "Gprop/s %.1f nodes/s %.0f ms/step %.1f frac %.4f e2e %.1f" % (d["value"] / 1e9, d["nodes_per_sec"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"] / 1e9))
PY
