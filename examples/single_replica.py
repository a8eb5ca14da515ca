"""One replica on the GPU, AsyncFlow-style (mirrors reference examples/yaml_input/single_server/single_server.py).

    python examples/single_replica.py tests/scenarios/c1_my_service.yml
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from asyncflow_b200 import GpuSimulationRunner  # noqa: E402

yaml_path = sys.argv[1] if len(sys.argv) > 1 else "tests/scenarios/c1_my_service.yml"
runner = GpuSimulationRunner.from_yaml(env=None, yaml_path=yaml_path, seed=42)
results = runner.run()                         # ResultsAnalyzer-compatible
print(results.format_latency_stats())
ts, rps = results.get_throughput_series()
print(f"throughput: mean {sum(rps) / len(rps):.1f} rps over {len(ts)} one-second buckets")
for sid in results.list_server_ids():
    t, ram = results.get_series("ram_in_use", sid)
    print(f"{sid}: RAM in use mean {sum(ram) / max(len(ram), 1):.1f} MB, max {max(ram, default=0)} MB")
try:                                           # the reference's own analyzer (and its plots), when installed
    analyzer = results.to_reference_analyzer()
    print("reference ResultsAnalyzer:", analyzer.get_latency_stats())
except ImportError:
    pass
