#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for i in 1 2 3; do
  /usr/bin/time -f "wall %e s" turbo_b200/bin/turbo -s -v tests/data/tiny.fzn 2>&1 | grep -E "start-up|solveTime|initTime|preprocessing_time|wall|solve_time" >> $O/startup.log
  echo "--" >> $O/startup.log
done
for i in 1 2; do
  /usr/bin/time -f "wall %e s" turbo_b200/bin/turbo -s -v -disable_simplify tests/data/tiny.fzn 2>&1 | grep -E "start-up|solveTime|initTime|preprocessing_time|wall" >> $O/startup.log
  echo "-- (no simplify)" >> $O/startup.log
done
TB_SNAPSHOT_MB=0 /usr/bin/time -f "wall %e s" turbo_b200/bin/turbo -s -v tests/data/tiny.fzn 2>&1 | grep -E "start-up|solveTime|wall" >> $O/startup.log
echo "-- (no snapshots)" >> $O/startup.log
python - <<'PY' >> gpurun_out/startup.log 2>&1
import time, sys
sys.path.insert(0, ".")
t = time.time()
from turbo_b200 import engine, abi
from tests import tnf_gen
pb = tnf_gen.planted(50, 60, 1)
print("import %.3f" % (time.time() - t))
for i in range(3):
    t = time.time(); s = engine.Solver(pb); t1 = time.time(); r = s.propagate(); t2 = time.time(); g = s.solve(); t3 = time.time(); s.close(); t4 = time.time()
    print("create %.3f propagate %.3f solve %.3f destroy %.3f  blocks %d" % (t1 - t, t2 - t1, t3 - t2, t4 - t3, r["stats"]["num_blocks"]))
PY
cat $O/startup.log
