#!/bin/bash
# Last check of a round on the final tree: smoke, the GPU test tier, the contract bench.
cd "$(dirname "$0")/.." || exit 1
tag=${1:-rXX}
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke
echo "=== pytest -m gpu"; timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "=== bench.py"; timeout 600 python bench.py > gpurun_out/bench_${tag}_final.json 2> gpurun_out/bench_${tag}_final.err; cut -c1-250 gpurun_out/bench_${tag}_final.json
python -c "
import json
d=json.loads(open('gpurun_out/bench_${tag}_final.json').readline()); print({k:d[k] for k in ('value','ms_per_step','passes','gpu_launches','clocks')}); print(d['e2e']); print(json.dumps(d['roofline'])[:1200]); print({k:d['cpu_baseline'][k] for k in ('value','cores','cores_effective','one_process_value','sample')})"
