#!/bin/bash
# Round-2 last GPU call (about 160 s of run time): A/B of the built engine libraries on the bench workload (device time +
# a checksum over every replica), the record-slot floor on the fastest one, then the lane-kernel GPU tests on that library.
#   gpurun --timeout 170 -- 'bash tools/gpu_ab_final.sh > gpurun_out/ab_r02k.log 2>&1'
cd "$(dirname "$0")/.." || exit 1
T0=$SECONDS
: > gpurun_out/ab_r02k.jsonl
for v in ${VARIANTS:-"" _pop4r _rqe _all}; do
  lib=$PWD/asyncflow_b200/_lib/libasyncflow_b200$v.so
  [ -f "$lib" ] || continue
  ASYNCFLOW_B200_LIB=$lib timeout 30 python tools/ab_lane_lib.py | tee -a gpurun_out/ab_r02k.jsonl
done
best=$(python - <<'PY'
import json
rows = [json.loads(l) for l in open("gpurun_out/ab_r02k.jsonl") if l.startswith("{")]
base = rows[0]
ok = [r for r in rows if r["checksum"] == base["checksum"] and r["flags"] == base["flags"]]
print(min(ok, key=lambda r: r["ms"])["lib"])
PY
)
echo "=== fastest with the product build's checksum: $best  (t=$((SECONDS-T0)) s)"
export ASYNCFLOW_B200_LIB=$best
for q in ${RQ_MINS:-3 2}; do ASYNCFLOW_B200_RQ_MIN=$q timeout 30 python tools/ab_lane_lib.py | tee -a gpurun_out/ab_r02k.jsonl; done
echo "=== GPU tests on $best  (t=$((SECONDS-T0)) s)"
left() { echo $(( ${BUDGET:-150} - (SECONDS - T0) )); }
# golden vectors first (no oracle run on the CPU: seconds), then the tests that re-run flagged replicas, compare launch
# shapes and pin the bench rows to the oracle, for as long as the budget lasts
[ $(left) -gt 15 ] && timeout $(left) python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider \
  -k "reproduces_golden_vectors and (lane or two_pass or auto)" > gpurun_out/ab_r02k_tests1.log 2>&1; tail -3 gpurun_out/ab_r02k_tests1.log
echo "=== (t=$((SECONDS-T0)) s)"
[ $(left) -gt 15 ] && timeout $(left) python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider \
  -k "flagged or picks_the_kernel or launch_shape or sweep_rows or bench_config or (pinned and (lane or auto))" > gpurun_out/ab_r02k_tests2.log 2>&1; tail -3 gpurun_out/ab_r02k_tests2.log
# configs[3]'s shape at 7 warps per SM (the 8-GPU run's occupancy): record-slot floor 4 against the former half of the even split
for q in 4 7; do [ $(left) -gt 14 ] && ASYNCFLOW_B200_RQ_MIN=$q timeout $(left) python tools/ab_lane_lib.py --config c4 --replicas 33152 --horizon 30 --reps 1 | tee -a gpurun_out/ab_r02k.jsonl; done
echo "=== done (t=$((SECONDS-T0)) s)"
