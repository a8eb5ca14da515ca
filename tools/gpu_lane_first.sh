#!/bin/bash
# First GPU session of round 2: the thread-per-replica engine -- device parity, then timings next to the
# warp-per-replica engine and its build variants.
#   gpurun --timeout 1500 -- 'bash tools/gpu_lane_first.sh > gpurun_out/lane_first.log 2>&1'
cd "$(dirname "$0")/.." || exit 1
QB="timeout 200 python tools/quick_bench.py"
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke || { echo "SMOKE FAILED: stopping"; exit 1; }
echo "=== parity, product library, mode auto"; timeout 400 python tools/check_variant_gpu.py || { echo "PARITY FAILED (auto): stopping"; exit 1; }
echo "=== parity, product library, mode lane"; ASYNCFLOW_B200_ENGINE=lane timeout 400 python tools/check_variant_gpu.py || echo "PARITY FAILED (lane)"
echo "=== parity, product library, mode warp"; ASYNCFLOW_B200_ENGINE=warp timeout 400 python tools/check_variant_gpu.py || echo "PARITY FAILED (warp)"
for mode in auto warp; do
  echo "=== C3 rtt sweep 40000 x 20 s, mode $mode"; $QB --scenario c3_lb_two_servers.yml --replicas 40000 --horizon 20 --reps 3 --mode $mode | tail -3
done
for wpb in 2 4 6 12 16; do
  echo "=== C3 rtt sweep, lane, $wpb warps/SM"; $QB --scenario c3_lb_two_servers.yml --replicas 40000 --horizon 20 --reps 2 --mode auto --wpb $wpb | tail -2
done
for mode in auto warp; do
  echo "=== C1 x 40000 x 60 s, mode $mode"; $QB --scenario c1_my_service.yml --replicas 40000 --horizon 60 --reps 2 --sweep none --mode $mode | tail -2
  echo "=== C2 users sweep 10000 x 60 s, mode $mode"; $QB --scenario c1_my_service.yml --replicas 10000 --horizon 60 --reps 2 --sweep users --mode $mode | tail -2
  echo "=== C4 20000 x 120 s, mode $mode"; $QB --scenario c4_lb8_events.yml --replicas 20000 --horizon 120 --reps 2 --sweep none --mode $mode | tail -2
  echo "=== C5 10000 x 30 s, mode $mode"; $QB --scenario c5_multihop32.yml --replicas 10000 --horizon 30 --reps 2 --sweep none --mode $mode | tail -2
done
echo "=== bench.py (auto)"; timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline | cut -c1-600
echo "=== bench.py (warp)"; ASYNCFLOW_B200_ENGINE=warp timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline | cut -c1-600
# round-1 build variants of the warp engine (libraries built in round 1: warp-per-replica only)
for v in _all _pin _all4; do
  lib="asyncflow_b200/_lib/libasyncflow_b200${v}.so"; [ -f "$lib" ] || continue
  echo "=== variant $v"
  ASYNCFLOW_B200_LIB="$PWD/$lib" timeout 200 python tools/quick_bench.py --scenario c3_lb_two_servers.yml --replicas 40000 --horizon 20 --reps 2 | tail -1
  ASYNCFLOW_B200_LIB="$PWD/$lib" timeout 200 python tools/quick_bench.py --scenario c1_my_service.yml --replicas 10000 --horizon 60 --reps 2 --sweep users | tail -1
  ASYNCFLOW_B200_LIB="$PWD/$lib" timeout 200 python tools/quick_bench.py --scenario c5_multihop32.yml --replicas 10000 --horizon 30 --reps 2 --sweep none | tail -1
done
