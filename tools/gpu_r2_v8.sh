#!/bin/bash
# Round 2, GPU visit 8: active set with the table in tensor memory and with direct marking (one barrier per sweep),
# the fixpoint without the publish fence, tail-splitting tests.
mkdir -p gpurun_out
O=gpurun_out
V=$PWD/turbo_b200/variants
( timeout -k 10 900 python -m pytest tests/test_gpu_active.py tests/test_gpu_configs.py tests/test_gpu_stream.py -m gpu -q --timeout 300 ) > $O/pytest_v8_default.log 2>&1; tail -4 $O/pytest_v8_default.log
for v in actdirect nofence; do
  ( TURBO_B200_LIB=$V/libturbo_b200_$v.so timeout -k 10 900 python -m pytest tests/test_gpu_active.py tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -q --timeout 300 ) > $O/pytest_v8_$v.log 2>&1
  echo "variant $v: $(tail -1 $O/pytest_v8_$v.log)"
done
B="--steps 5 --warmup 3 --no-cpu-baseline --strong-ms 0 --e2e-steps 2 --no-fixpoint-leg"
for v in default actdirect; do
  if [ $v = default ]; then unset TURBO_B200_LIB; else export TURBO_B200_LIB=$V/libturbo_b200_$v.so; fi
  for w in simplified:trains15 simplified:example_wordpress7_500 trains15; do
    timeout 300 python bench.py $B --fp wac1_active --workload $w > $O/ab8_${v}_active_$(echo $w | tr ':' '_').json 2>> $O/ab8.err
  done
done
unset TURBO_B200_LIB
TB_TMEM=0 timeout 300 python bench.py $B --fp wac1_active > $O/ab8_default_notmem_active_simplified_trains15.json 2>> $O/ab8.err
for v in default nofence; do
  if [ $v = default ]; then unset TURBO_B200_LIB; else export TURBO_B200_LIB=$V/libturbo_b200_$v.so; fi
  timeout 300 python bench.py $B > $O/ab8_${v}_dense_trains15.json 2>> $O/ab8.err
  timeout 300 python bench.py $B --workload simplified:accap_a3 > $O/ab8_${v}_dense_accap.json 2>> $O/ab8.err
  timeout 300 python tools/fixpoint_bench.py --workload synthetic:100000:1000000 --mem store_cluster --repeat 5 --rounds 2 > $O/ab8_${v}_synthetic_cluster.json 2>> $O/ab8.err
  timeout 300 python tools/fixpoint_bench.py --workload trains15 --repeat 20 --rounds 3 > $O/ab8_${v}_fixpoint_trains15.json 2>> $O/ab8.err
done
unset TURBO_B200_LIB
for f in $O/ab8_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    if "value" in d:
        print(sys.argv[1].split("/")[-1], d["config"]["memory_configuration"], d["config"]["fixpoint"], "Gprop/s %.1f nodes/s %.0f fpshare %.2f" % (d["value"]/1e9, d["nodes_per_sec"], d["fixpoint_time_share"] or 0))
    else:
        print(sys.argv[1].split("/")[-1], d["memory_configuration"], "Gprop/s %.1f frac %.3f" % (d["propagations_per_sec"]/1e9, d["smem_frac"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
done
