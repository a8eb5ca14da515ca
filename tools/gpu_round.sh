#!/bin/bash
# Round evidence in one GPU call: the GPU test tier, the contract bench (both arms), the other BASELINE shapes on one GPU.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r02 > gpurun_out/round_r02.log 2>&1'
tag=${1:-rXX}
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "=== bench.py (contract defaults)"; timeout 600 python bench.py > gpurun_out/bench_${tag}_c3.json 2> gpurun_out/bench_${tag}_c3.err; cut -c1-300 gpurun_out/bench_${tag}_c3.json
echo "=== bench.py --impl reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_c3_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_${tag}_c3_reference.json
for cfg in ${CONFIGS:-"c2 0 0" "c4 40000 250" "c5 40000 10"}; do
  set -- $cfg
  extra=""; [ "$2" != "0" ] && extra="--replicas $2 --horizon $3"
  echo "=== bench.py --config $1 $extra (one GPU)"
  timeout 900 python bench.py --config $1 $extra --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_${tag}_$1_1gpu.json 2> gpurun_out/bench_${tag}_$1.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_${tag}_$1_1gpu.json').readline()); print({k:d[k] for k in ('value','ms_per_step','replicas_overflowed','passes','events_per_s')}, 'e2e', d['e2e']['value'])"
done
