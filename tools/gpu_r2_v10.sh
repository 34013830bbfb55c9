#!/bin/bash
# Round 2, GPU visit 10: whole GPU suite after the co-residency fix, the default bench + reference arm, ncu of the final
# solve kernel (all-in-TMEM variant), DRAM traffic of a default launch, compute-sanitizer on a STORE_CLUSTER fixpoint.
mkdir -p gpurun_out
O=gpurun_out
( time timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 300 --durations=6 ) > $O/pytest_gpu_v10.log 2>&1; tail -14 $O/pytest_gpu_v10.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > $O/bench_default_v10.json 2> $O/bench_default_v10.err; head -c 2200 $O/bench_default_v10.json; echo
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm_v10.json 2>> $O/bench_default_v10.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_v10.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --strong-ms 0 --e2e-steps 1 > $O/ncu_launches_v10.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -f -o $O/solve_final_trains15 \
  python bench.py --steps 1 --warmup 0 --cutnodes 300 --no-cpu-baseline --no-fixpoint-leg --strong-ms 0 --e2e-steps 0 > $O/ncu_solve_final.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,lts__t_bytes.sum --clock-control none -k regex:solve_kernel -c 1 --csv --log-file $O/traffic_default_launch_v10.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-fixpoint-leg --strong-ms 0 --e2e-steps 0 > $O/ncu_traffic_v10.log 2>&1
grep solve_kernel $O/traffic_default_launch_v10.csv | cut -d, -f13- | head -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:propagate_kernel -c 1 -f -o $O/propagate_final_trains15 \
  python tools/fixpoint_bench.py --workload trains15 --repeat 20 --rounds 0 > $O/ncu_fixpoint_final.log 2>&1
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python - <<'P'
import numpy as np
from tests import tnf_gen
from turbo_b200 import abi, engine
from oracle import oracle_py as orc
for cs in (4, 16):
    pb = tnf_gen.planted(5000, 20000, 71)
    o = orc.fixpoint(pb)
    with engine.Solver(pb, mem_kind=abi.MEM_STORE_CLUSTER, cluster_size=cs) as s:
        g = s.propagate()
    assert np.array_equal(g["lb"], o["lb"]) and np.array_equal(g["ub"], o["ub"]), cs
    print("cluster", cs, "fixpoint bit-exact under memcheck")
pb = tnf_gen.search_instance(3)
with engine.Solver(pb, mem_kind=abi.MEM_STORE_CLUSTER, cluster_size=2, subproblems_power=3) as s:
    r = s.solve()
print("cluster solve under memcheck: exhaustive", r["exhaustive"], "objective", r["objective"])
P
) > $O/sanitizer_cluster.log 2>&1; echo "compute-sanitizer rc $?"; tail -6 $O/sanitizer_cluster.log
