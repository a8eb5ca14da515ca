"""One engine build on the bench workload, without torch: device time per launch and a checksum of every replica.

    ASYNCFLOW_B200_LIB=asyncflow_b200/_lib/libasyncflow_b200_x.so python tools/ab_lane_lib.py [--config c3] [--reps 2]

Prints ONE JSON line: ms per launch (best of the repetitions, device-timed), completions/s, the pass structure, and
`checksum` = exact sums over all replicas (completions, events, fsum of the latency sums and of their squares): two
builds that simulate the same replicas bit for bit print the same checksum.  Session tooling (tools/gpu_ab_final.sh)."""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402

from asyncflow_b200 import SweepRunner, flatten  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--replicas", type=int, default=None)
    ap.add_argument("--horizon", type=int, default=None)
    a = ap.parse_args()
    w = bench.make_workload(a.config, a.horizon, a.replicas)
    n = w.replicas
    flat = flatten(w.payload)
    sw = SweepRunner(flat, n, w.columns(flat, np.arange(n, dtype=np.int64), n), seed=bench.SEED, histogram=True,
                     throughput=False, pinned=False)
    eng = sw.engine()
    eng.upload_sweep(sw.spec, 0, row_first=0, row_count=n)
    ms = []
    for _ in range(1 + a.reps):                     # the first launch allocates
        eng.configure(request_capacity=sw.request_capacity, event_capacity=sw.event_capacity, histogram=True, throughput=False)
        eng.run(bench.SEED, 0, n)
        eng.sync()
        ms.append(eng.last_run_ms()[0])
    st = eng.stats()
    p = eng.last_run_passes()
    best = min(ms[1:])
    print(json.dumps({
        "lib": os.environ.get("ASYNCFLOW_B200_LIB", "product"), "rq_min": os.environ.get("ASYNCFLOW_B200_RQ_MIN"),
        "config": a.config, "ms": best, "ms_all": ms, "value": float(st["completed"].sum()) / (best / 1e3),
        "events_smem": p["lane_events_smem"], "requests_smem": p["lane_requests_smem"], "warps": p["lane_warps_per_sm"],
        "rerun": p["warp_replicas"] if p["lane_pass"] else None, "flags": int((st["flags"] != 0).sum()),
        "checksum": [int(st["completed"].sum()), int(st["n_events"].sum()), math.fsum(st["lat_sum"]).hex(),
                     math.fsum(st["lat_sumsq"]).hex(), int(st["n_ticks"].sum())]}))


if __name__ == "__main__":
    main()
