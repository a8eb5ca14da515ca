"""A small workload for compute-sanitizer (tools/sanitize_gpu.sh): C1, C5 and the overloaded single server at 64
replicas each, through the C ABI, in the pass structure named by ASYNCFLOW_B200_ENGINE (default auto)."""
from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import yaml  # noqa: E402

from asyncflow_b200 import Engine, flatten  # noqa: E402

CASES = (("c1_my_service.yml", 5), ("c5_multihop32.yml", 2), ("overload_single.yml", 6), ("tie_cpu_io.yml", 5),
         ("ev_spikes_outages.yml", None))
with Engine(0) as eng:
    for name, horizon in CASES:
        d = yaml.safe_load((ROOT / "tests" / "scenarios" / name).read_text())
        if horizon:
            d["sim_settings"]["total_simulation_time"] = horizon
        flat = flatten(d)
        eng.upload(flat)
        eng.configure(trace_replicas=2, trace_clock_capacity=20000, throughput=True)
        eng.run(0xA5F10, 0, 64)
        st = eng.stats()
        eng.reduced_histogram()
        print(name, "completed", int(st["completed"].sum()), "flags", sorted(set(int(f) for f in st["flags"])), eng.last_run_passes())
