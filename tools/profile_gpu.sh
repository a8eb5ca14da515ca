#!/bin/bash
# The round's ncu passes in one GPU call (never under torchrun; numbers printed under ncu are not bench values):
#   gpurun --timeout 1500 -- 'bash tools/profile_gpu.sh r02a > gpurun_out/profile_r02a.log 2>&1'
# then, here:  python tools/ncu_issue_summary.py r02a          (writes profiles/kernel_metrics.json + copies the CSVs)
#              ncu -i gpurun_out/prof_<tag>.ncu-rep --page source --csv --print-source sass > /tmp/sass.csv
#              python tools/ncu_by_function.py /tmp/sass.csv --trim profiles/<tag>_ncu_sass_executed.csv > profiles/<tag>_instructions_by_function.txt
tag=${1:-rXX}
kernel=${2:-af_lane_kernel}
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
# 1. launch list of the bench command (kernel shares of the step; per-launch times are cold-cache and serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_launches_${tag}.log 2>&1
# 2. ONE bench-size launch of the dominant kernel: DRAM bytes (roofline.traffic), executed warp instructions, issue-slot
#    utilisation, stall reasons per issued instruction (roofline.issue) -- a metrics list, seconds, not `--set full`
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum
M=$M,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
M=$M,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio
M=$M,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:${kernel} -c 1 --csv --log-file gpurun_out/metrics_${tag}.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/b_metrics_${tag}.log 2>&1
# 3. full capture with source, on a launch small enough for ~40 replays
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${kernel} -c 1 -o gpurun_out/prof_${tag} \
    python bench.py --steps 1 --warmup 0 --horizon 5 --replicas 38000 --no-cpu-baseline > gpurun_out/b_full_${tag}.log 2>&1
ls -la gpurun_out | tail -8
