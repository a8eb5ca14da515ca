#!/bin/bash
# compute-sanitizer over the product library, small grids:
#   gpurun --timeout 1200 -- 'bash tools/sanitize_gpu.sh r02 > gpurun_out/sanitize_r02.log 2>&1'
# memcheck for every pass structure; racecheck (shared-memory hazards) on the two kernels that use shared memory.
tag=${1:-rXX}
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
for mode in two_pass lane warp; do
  echo "=== memcheck, mode $mode"
  ASYNCFLOW_B200_ENGINE=$mode timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/sanitize_${tag}_memcheck_${mode}.log 2>&1
  echo "exit $?"; tail -4 gpurun_out/sanitize_${tag}_memcheck_${mode}.log
done
for mode in lane warp; do
  echo "=== racecheck, mode $mode"
  ASYNCFLOW_B200_ENGINE=$mode timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/sanitize_${tag}_racecheck_${mode}.log 2>&1
  echo "exit $?"; tail -4 gpurun_out/sanitize_${tag}_racecheck_${mode}.log
done
