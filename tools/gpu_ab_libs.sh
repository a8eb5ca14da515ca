#!/bin/bash
# A/B of engine builds in ONE GPU call: for every asyncflow_b200/_lib/libasyncflow_b200*.so (ASYNCFLOW_B200_LIB selects it)
# the bench line, then the issue / fetch-stall metrics of one small launch.
#   gpurun --timeout 900 -- 'bash tools/gpu_ab_libs.sh > gpurun_out/ab_TAG.log 2>&1'
cd "$(dirname "$0")/.." || exit 1
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
M=$M,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
timeout 200 python __graft_entry__.py --smoke || exit 1
for lib in ${LIBS:-$(ls asyncflow_b200/_lib/libasyncflow_b200*.so)}; do
  echo "=== $lib"
  export ASYNCFLOW_B200_LIB="$PWD/$lib"
  timeout 200 python tools/check_parity_gpu.py 2>&1 | tail -1
  for wpb in ${WPBS:-0}; do
    timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --wpb $wpb | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('bench wpb $wpb', {k:d[k] for k in ('value','ms_per_step','replicas_overflowed')}, d['passes']['lane_warps_per_sm'], d['passes']['lane_events_in_smem'])"
  done
  timeout 200 python tools/quick_bench.py --scenario c1_my_service.yml --replicas 60000 --horizon 60 --reps 2 --sweep none | grep run1 | cut -c1-200
  timeout 300 ncu --metrics $M --clock-control none -k regex:af_lane_kernel -c 1 --csv --log-file /tmp/m.csv python bench.py --steps 1 --warmup 0 --horizon 10 --replicas 38000 --no-cpu-baseline > /dev/null 2>&1
  grep -v "^==" /tmp/m.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"' | grep -v "Metric Name" | sed 's/smsp__average_warps_issue_stalled_//; s/_per_issue_active.ratio//' | tr '\n' ';'; echo
done
