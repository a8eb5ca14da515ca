"""SURVEY §8d's CPU-side detail: the reference's execution model on ONE core, all cores, and the bare
heap loop of the simpy-compatible kernel the oracle runs on.  Test/bench tooling (uses oracle/).

    python tools/cpu_baseline_detail.py [--horizon 60]
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "tests", ROOT / "oracle", ROOT / "oracle" / "simpy_shim"):
    sys.path.insert(0, str(p))


def one(args):
    import des_port
    from helpers import SEED, load_scenario
    name, horizon, replica = args
    o = des_port.simulate(load_scenario(name, horizon), seed=SEED, replica=replica)
    return len(o["clocks"]), int(o.get("heap_events", 0))


def heap_loop(n: int) -> float:
    """n timeouts through the shim's Environment.step(): one generator process, nothing else."""
    import simpy
    env = simpy.Environment()

    def proc():
        for _ in range(n):
            yield env.timeout(1.0)
    env.process(proc())
    t0 = time.perf_counter()
    env.run()
    return n / (time.perf_counter() - t0)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--horizon", type=int, default=60)
    a = ap.parse_args()
    out = {"cores": os.cpu_count()}
    for name in ("c1_my_service.yml", "c3_lb_two_servers.yml"):
        one((name, 5, 0))
        t0 = time.perf_counter()
        n, ev = one((name, a.horizon, 1))
        dt = time.perf_counter() - t0
        out[name] = {"one_process": {"completions_per_s": n / dt, "heap_events_per_s": ev / dt, "wall_s": dt}}
        k = os.cpu_count() or 1
        with mp.get_context("fork").Pool(k) as pool:
            pool.map(one, [(name, 2, i) for i in range(k)])
            t0 = time.perf_counter()
            res = pool.map(one, [(name, a.horizon, i) for i in range(2 * k)])
            dt = time.perf_counter() - t0
        out[name]["all_cores"] = {"processes": k, "completions_per_s": sum(r[0] for r in res) / dt,
                                  "heap_events_per_s": sum(r[1] for r in res) / dt, "wall_s": dt}
    out["shim_heap_loop_events_per_s"] = heap_loop(300_000)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
