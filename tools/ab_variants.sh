#!/bin/bash
# A/B the engine's build variants against the product build in ONE GPU session:
#   gpurun --timeout 2400 -- 'bash tools/ab_variants.sh > gpurun_out/ab.log 2>&1'      (~3 GPU-min per library)
# For every library: device parity vs the oracle first, then the same three timings
# (C3 as written; C1; saturated C2; C4; bench.py).  Build the variants beforehand on the
# build box (python tools/build_variants.py) -- the .so files travel with the snapshot.
cd "$(dirname "$0")/.." || exit 1
# VARIANTS="_all _sorted" bash tools/ab_variants.sh   picks a subset ("" = the product build, always first)
for v in "" ${VARIANTS:-_predraw _pregen _memo _sorted _pin _all _all4 _all_mb6}; do
  lib="asyncflow_b200/_lib/libasyncflow_b200${v}.so"
  [ -f "$lib" ] || { echo "missing $lib"; continue; }
  echo "=== $lib"
  export ASYNCFLOW_B200_LIB="$PWD/$lib"
  timeout 300 python tools/check_variant_gpu.py || { echo "PARITY FAILED for $lib"; continue; }
  timeout 200 python tools/quick_bench.py --scenario c3_lb_two_servers.yml --replicas 40000 --horizon 20 --reps 3 | tail -2
  timeout 200 python tools/quick_bench.py --scenario c1_my_service.yml --replicas 40000 --horizon 60 --reps 2 --sweep none | tail -1
  timeout 200 python tools/quick_bench.py --scenario c1_my_service.yml --replicas 10000 --horizon 60 --reps 2 --sweep users | tail -1
  timeout 200 python tools/quick_bench.py --scenario c4_lb8_events.yml --replicas 20000 --horizon 120 --reps 2 --sweep none | tail -1
  echo "--- saturated C2 again, heaviest rows first (host-side launch order)"
  timeout 200 python tools/quick_bench.py --scenario c1_my_service.yml --replicas 10000 --horizon 60 --reps 2 --sweep users --balance | tail -1
  timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline | cut -c1-400
done
