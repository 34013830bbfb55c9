#!/bin/bash
python - <<'PY'
import sys, time, os
sys.path.insert(0, ".")
from tests import golden_io
from turbo_b200 import abi, engine
pb, info = golden_io.load_simplified_problem("trains15")
for mb in ("4096", "0"):
    os.environ["TB_SNAPSHOT_MB"] = mb
    for i in range(4):
        t0 = time.perf_counter(); s = engine.Solver(pb, cutnodes=2000); t1 = time.perf_counter(); r = s.solve(); t2 = time.perf_counter(); s.close(); t3 = time.perf_counter()
        print("snapshot MB %s: create %.2f ms  solve %.2f ms (kernel %.2f)  destroy %.2f ms" % (mb, (t1 - t0) * 1e3, (t2 - t1) * 1e3, r["stats"]["kernel_ms"], (t3 - t2) * 1e3))
PY
