#!/bin/bash
# A/B of the lane's shared-memory split between pending events and request records (host-side policy, one GPU call):
# ASYNCFLOW_B200_EVEN_SPLIT=1 = equal counts (round-2 builds up to r2i), default = what the replica typically needs.
cd "$(dirname "$0")/.." || exit 1
timeout 200 python __graft_entry__.py --smoke || exit 1
timeout 300 python tools/check_parity_gpu.py 2>&1 | tail -1
for even in 1 0; do
  echo "=== EVEN_SPLIT=$even"
  [ "$even" = "1" ] && export ASYNCFLOW_B200_EVEN_SPLIT=1 || unset ASYNCFLOW_B200_EVEN_SPLIT
  timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('bench', {k:d[k] for k in ('value','ms_per_step','replicas_overflowed')}, d['passes']['lane_warps_per_sm'], d['passes']['lane_events_in_smem'], d['passes']['lane_requests_in_smem'])"
  timeout 200 python tools/quick_bench.py --scenario c1_my_service.yml --replicas 60000 --horizon 60 --reps 2 --sweep none | grep -E "run1|passes" | cut -c1-260
  timeout 200 python tools/quick_bench.py --scenario c4_lb8_events.yml --replicas 40000 --horizon 120 --reps 1 --sweep none | grep -E "run0|passes" | cut -c1-260
  timeout 200 python tools/quick_bench.py --scenario c5_multihop32.yml --replicas 40000 --horizon 8 --reps 1 --sweep none --mode two_pass | grep -E "run0|passes" | cut -c1-260
done
