#!/bin/bash
# configs[3] and configs[4] at BASELINE size on 8 GPUs, one gpurun --gpus 8 call:
#   gpurun --gpus 8 --timeout 540 -- 'bash tools/gpu_8gpu_configs.sh r02 > gpurun_out/8gpu_r02.log 2>&1'
tag=${1:-rXX}
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for cfg in ${CONFIGS:-"c4 250" "c5 8"}; do
  set -- $cfg
  echo "=== bench.py --gpus 8 --config $1 --horizon $2 (BASELINE replica count)"
  timeout ${PER_BENCH:-260} python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 8 --config $1 --horizon $2 --steps 1 --warmup 1 --e2e-warm 0 > gpurun_out/bench_${tag}_$1_8gpu.json 2> gpurun_out/bench_${tag}_$1_8gpu.err
  echo "exit $?"
  python -c "
import json
d=json.loads(open('gpurun_out/bench_${tag}_$1_8gpu.json').readline()); print({k:d[k] for k in ('value','n_gpus','ms_per_step','per_rank_ms','replicas_overflowed','passes','events_per_s')}, 'config', d['config']['replicas_total'], d['config']['horizon_s'], 'e2e', d['e2e'])" || tail -5 gpurun_out/bench_${tag}_$1_8gpu.err
done
