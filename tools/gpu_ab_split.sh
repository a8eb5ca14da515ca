#!/bin/bash
# A/B of the lane's shared-memory split between pending events and request records (host-side knob, one GPU call).
cd "$(dirname "$0")/.." || exit 1
timeout 200 python __graft_entry__.py --smoke || exit 1
for share in ${SHARES:-0 60 70 80 88}; do
  echo "=== ASYNCFLOW_B200_EV_SHARE=$share"
  ASYNCFLOW_B200_EV_SHARE=$share timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('bench', {k:d[k] for k in ('value','ms_per_step','replicas_overflowed')}, d['passes']['lane_warps_per_sm'], d['passes']['lane_events_in_smem'], d['passes']['lane_requests_in_smem'])"
  ASYNCFLOW_B200_EV_SHARE=$share timeout 200 python tools/quick_bench.py --scenario c1_my_service.yml --replicas 60000 --horizon 60 --reps 2 --sweep none | grep -E "run1|passes" | cut -c1-260
  ASYNCFLOW_B200_EV_SHARE=$share timeout 200 python tools/quick_bench.py --scenario c4_lb8_events.yml --replicas 40000 --horizon 120 --reps 1 --sweep none | grep -E "run0|passes" | cut -c1-260
done
