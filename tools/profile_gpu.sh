#!/bin/bash
# The round's three ncu passes in one GPU call (never under torchrun; numbers printed under ncu are not bench values):
#   gpurun --timeout 1500 -- 'bash tools/profile_gpu.sh r02a > gpurun_out/profile_r02a.log 2>&1'
# then, here:  ncu -i gpurun_out/prof_<tag>.ncu-rep --page source --csv --print-source sass > /tmp/sass.csv
#              python tools/ncu_by_function.py /tmp/sass.csv --trim profiles/<tag>_ncu_sass_executed.csv > profiles/<tag>_instructions_by_function.txt
# and copy the CSVs below into profiles/ (see profiles/r01_summary.md for what each one is).
tag=${1:-rXX}
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
# 1. launch list of the bench command (kernel shares of the step; per-launch times are cold-cache and serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_launches_${tag}.log 2>&1
# 2. DRAM bytes of ONE bench-size launch of the sim kernel (roofline.traffic)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:af_sim_kernel -c 1 --csv --log-file gpurun_out/dram_${tag}.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/b_dram_${tag}.log 2>&1
# 3. full capture with source, on a launch small enough for ~40 replays (20 000 replicas x 5 s: ~1 min)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:af_sim_kernel -c 1 -o gpurun_out/prof_${tag} \
    python bench.py --steps 1 --warmup 0 --horizon 5 --replicas 20000 --no-cpu-baseline > gpurun_out/b_full_${tag}.log 2>&1
ls -la gpurun_out | tail -8
