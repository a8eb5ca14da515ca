cd "$(dirname "$0")/.." || exit 1
for half in 1 0; do
  [ "$half" = "1" ] && export ASYNCFLOW_B200_RQ_MIN_HALF=1 || unset ASYNCFLOW_B200_RQ_MIN_HALF
  echo "=== RQ_MIN_HALF=$half"
  timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('c3', {k:d[k] for k in ('value','ms_per_step')}, d['passes']['lane_events_in_smem'], d['passes']['lane_requests_in_smem'])"
  timeout 200 python bench.py --config c4 --replicas 66304 --horizon 120 --steps 1 --warmup 1 --no-cpu-baseline --e2e-warm 0 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('c4', {k:d[k] for k in ('value','ms_per_step')}, d['passes']['lane_warps_per_sm'], d['passes']['lane_events_in_smem'], d['passes']['lane_requests_in_smem'])"
done
