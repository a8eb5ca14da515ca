#!/bin/bash
# One iteration of the lane-engine work on the GPU: parity first, then the timings that decide, then a light ncu pass.
#   gpurun --timeout 900 -- 'bash tools/gpu_lane_iter.sh TAG > gpurun_out/iter_TAG.log 2>&1'
tag=${1:-x}
cd "$(dirname "$0")/.." || exit 1
QB="timeout 200 python tools/quick_bench.py"
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke || { echo "SMOKE FAILED: stopping"; exit 1; }
echo "=== parity (auto)"; timeout 400 python tools/check_parity_gpu.py || { echo "PARITY FAILED (auto): stopping"; exit 1; }
for wpb in ${WPBS:-12}; do
  echo "=== C3 rtt sweep 80000 x 20 s, lane, $wpb warps/SM"; $QB --scenario c3_lb_two_servers.yml --replicas 80000 --horizon 20 --reps 2 --mode auto --wpb $wpb | tail -2
done
echo "=== C1 x 40000 x 60 s"; $QB --scenario c1_my_service.yml --replicas 40000 --horizon 60 --reps 2 --sweep none | tail -2
echo "=== C4 20000 x 120 s"; $QB --scenario c4_lb8_events.yml --replicas 20000 --horizon 120 --reps 2 --sweep none | tail -2
echo "=== C2 users sweep 10000 x 60 s"; $QB --scenario c1_my_service.yml --replicas 10000 --horizon 60 --reps 1 --sweep users | tail -2
echo "=== C2 users sweep 10000 x 60 s, no page pool (flagged replicas re-run per warp)"; ASYNCFLOW_B200_NO_PAGES=1 $QB --scenario c1_my_service.yml --replicas 10000 --horizon 60 --reps 1 --sweep users | tail -2
echo "=== C5 10000 x 10 s"; $QB --scenario c5_multihop32.yml --replicas 10000 --horizon 10 --reps 1 --sweep none | tail -2
for wpb in ${BENCH_WPBS:-0}; do
  echo "=== bench.py --wpb $wpb"; timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --wpb $wpb | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print({k:d[k] for k in ('value','ms_per_step','replicas_overflowed','passes')}, d['e2e']['value'])"
done
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
M=$M,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum
echo "=== ncu metrics, bench workload at 38000 x 10 s"
timeout 300 ncu --metrics $M --clock-control none -k regex:af_lane_kernel -c 1 --csv --log-file gpurun_out/metrics_${tag}.csv \
    python bench.py --steps 1 --warmup 0 --horizon 10 --replicas 38000 --no-cpu-baseline > gpurun_out/b_metrics_${tag}.log 2>&1
grep -v "^==" gpurun_out/metrics_${tag}.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"'
if [ -n "$FULL" ]; then
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:af_lane_kernel -c 1 -o gpurun_out/prof_${tag} \
      python bench.py --steps 1 --warmup 0 --horizon 5 --replicas 38000 --no-cpu-baseline > gpurun_out/b_full_${tag}.log 2>&1
fi
