#!/bin/bash
# Round 2, GPU visit 9: pat13 with and without tail splitting (a 120 s run did not finish in visit 8), the all-in-TMEM
# kernel variant, whole GPU suite.
mkdir -p gpurun_out
O=gpurun_out
cat > /tmp/pat13.py <<'P'
import sys, time, json
sys.path.insert(0, ".")
from tests import golden_io
from turbo_b200 import engine, abi
for name in ("pat13", "pat12"):
    pb, info = golden_io.load(name)
    for rep in range(3):
        t = time.time()
        with engine.Solver(pb, timeout_ms=40000) as s:
            cfg = s.config(); r = s.solve()
        st = r["stats"]
        print(json.dumps({"name": name, "rep": rep, "env": {k: __import__("os").environ.get(k) for k in ("TB_SPLIT_BITS",)}, "mem": cfg["mem_kind"], "blocks": cfg["num_blocks"], "tpb": cfg["threads_per_block"], "sub": cfg["subproblems_power"],
                          "exhaustive": r["exhaustive"], "obj": golden_io.user_objective(info, r["lb"], r["ub"]) if r["has_solution"] else None, "expected": info["expected"],
                          "secs": round(time.time() - t, 2), "nodes": st["nodes"], "solved": st["eps_solved_subproblems"], "skipped": st["eps_skipped_subproblems"],
                          "split": st["eps_split_subproblems"], "parts": st["eps_split_parts_solved"], "done": st["num_blocks_done"]}), flush=True)
P
python /tmp/pat13.py > $O/v9_pat13_split.jsonl 2>&1
TB_SPLIT_BITS=0 python /tmp/pat13.py > $O/v9_pat13_nosplit.jsonl 2>&1
cat $O/v9_pat13_split.jsonl $O/v9_pat13_nosplit.jsonl | cut -c1-400
B="--steps 5 --warmup 3 --no-cpu-baseline --strong-ms 0 --e2e-steps 3"
timeout 300 python bench.py $B > $O/ab9_tmall_trains15.json 2> $O/ab9.err
timeout 300 python bench.py $B --workload trains15 --no-fixpoint-leg > $O/ab9_tmall_trains15full.json 2>> $O/ab9.err
timeout 300 python bench.py $B --workload simplified:example_wordpress7_500 --no-fixpoint-leg > $O/ab9_mixed_wordpress.json 2>> $O/ab9.err
for f in $O/ab9_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    fk=d.get("fixpoint_kernel",{}); a=d.get("active_set",{})
    print(sys.argv[1].split("/")[-1], "blocks %d Gprop/s %.1f nodes/s %.0f frac %.3f fpshare %.2f e2e %.1f | fixpoint-alone %.1f | active nodes/s %.0f" % (d["config"]["num_blocks_per_gpu"], d["value"]/1e9, d["nodes_per_sec"], d["roofline"]["frac"], d["fixpoint_time_share"] or 0, d["e2e"]["value"]/1e9, fk.get("propagations_per_sec",0)/1e9, a.get("nodes_per_sec",0)))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
done
( time timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 300 --durations=8 ) > $O/pytest_gpu_v9.log 2>&1; tail -16 $O/pytest_gpu_v9.log
