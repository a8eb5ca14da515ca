"""Parity of ONE engine library (product or variant) on the GPU: golden vectors + fuzz vs the oracle.

    ASYNCFLOW_B200_LIB=asyncflow_b200/_lib/libasyncflow_b200_memo.so python tools/check_parity_gpu.py

Test tooling for A/B sessions (uses oracle/); exits non-zero on the first mismatch.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "tests", ROOT / "oracle", ROOT / "oracle" / "simpy_shim"):
    sys.path.insert(0, str(p))

import des_port  # noqa: E402
import fuzz  # noqa: E402
from helpers import PARITY_CASES, SEED, assert_matches_oracle, load_scenario  # noqa: E402

from asyncflow_b200 import Engine, flatten  # noqa: E402


def check(eng: Engine, payload: dict, replica: int, event_capacity: int = 0) -> int:
    flat = flatten(payload)
    eng.upload(flat)
    eng.configure(trace_replicas=1, trace_clock_capacity=400000, request_capacity=400000, throughput=True,
                  event_capacity=event_capacity)
    eng.run(SEED, replica, replica + 1)
    st = eng.stats()
    sent, dropped = eng.edge_counts()
    assert st[0]["flags"] == 0, int(st[0]["flags"])
    o = des_port.simulate(payload, seed=SEED, replica=replica)
    assert_matches_oracle(o, flat, stats=st[0], clocks=eng.trace_clocks(0), sent=sent[0], dropped=dropped[0],
                          series=eng.trace_series(0), throughput=eng.throughput()[0])
    return int(st[0]["completed"])


def main() -> None:
    print("library:", os.environ.get("ASYNCFLOW_B200_LIB", "(product build)"))
    total = 0
    with Engine(0) as eng:
        for name, horizon in PARITY_CASES.items():
            total += check(eng, load_scenario(name, horizon), 7)
        for seed in range(500, 540):
            total += check(eng, fuzz.scenario(seed), seed)
        for seed in range(40, 48):                       # C5-shaped, overloaded: both HBM tiers, unsorted pool
            total += check(eng, fuzz.big_scenario(seed), seed, event_capacity=8192)
    print(f"OK: {len(PARITY_CASES)} scenarios + 40 random + 8 big random scenarios, {total} completions bit-exact vs the oracle")


if __name__ == "__main__":
    main()
