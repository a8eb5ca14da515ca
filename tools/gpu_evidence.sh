#!/bin/bash
# The round's ncu evidence + compute-sanitizer + the GPU test tier in one call:
#   gpurun --timeout 1200 -- 'bash tools/gpu_evidence.sh r02 > gpurun_out/evidence_r02.log 2>&1'
tag=${1:-rXX}
cd "$(dirname "$0")/.." || exit 1
echo "=== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "=== ncu passes"; bash tools/profile_gpu.sh $tag af_lane_kernel 2>&1 | tail -12
echo "=== compute-sanitizer"; bash tools/sanitize_gpu.sh $tag 2>&1
