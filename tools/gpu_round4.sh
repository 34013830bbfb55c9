#!/bin/bash
# GPU visit 4: new simplified fixtures (functional elimination) — tests, benches, ncu capture of the TCN_SHARED solve.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_simplifier.py tests/test_cli_gpu.py -m gpu -x -q ) > $O/pytest_simp2.log 2>&1
tail -3 $O/pytest_simp2.log
timeout 600 python bench.py --workload simplified:trains15 > $O/bench_strains2.json 2> $O/bench_strains2.err
timeout 300 python bench.py --workload simplified:trains15 --no-cpu-baseline --no-fixpoint-leg --tpb 512 > $O/bench_strains2_512.json 2>> $O/bench_strains2.err
timeout 300 python bench.py --workload trains15 --no-cpu-baseline > $O/bench_trains_again.json 2>> $O/bench_strains2.err
for w in simplified:trains15; do
  timeout 60 python tools/time_to_optimum.py $w --timeout-ms 20000 >> $O/tto_simplified2.jsonl 2>> $O/tto.err
done
cat $O/tto_simplified2.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -f -o $O/solve_tcn_strains \
  python bench.py --workload simplified:trains15 --steps 1 --warmup 0 --cutnodes 300 --no-cpu-baseline --no-fixpoint-leg > $O/ncu_solve_tcn.log 2>&1
tail -2 $O/ncu_solve_tcn.log
python - <<'PY'
import json
for f in ("bench_strains2", "bench_strains2_512", "bench_trains_again"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        fk = d.get("fixpoint_kernel", {})
        print(f, d["config"]["memory_configuration"], d["config"]["threads_per_block"], d["config"]["num_blocks_per_gpu"],
              "Gprop/s %.1f nodes/s %.0f frac %.4f e2e %.1f" % (d["value"] / 1e9, d["nodes_per_sec"], d["roofline"]["frac"], d["e2e"]["value"] / 1e9))
    except Exception as e:
        print(f, "ERR", e)
PY
