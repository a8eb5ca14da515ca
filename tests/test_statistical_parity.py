"""Statistical parity with the reference run on ITS OWN RNG (numpy PCG64, libm log).

tests/golden/reference_numpy_stats.json holds, per scenario, the mean over ~100 replicas (and its
standard error) of what ``ResultsAnalyzer`` reports when the unmodified reference runs exactly as
upstream runs it (oracle/make_reference_stats.py).  The engine draws different random numbers
(AF-RNG), so the comparison is distributional: the sweep mean of every statistic must agree with
the reference's within 2 % (BASELINE.json: "p50/p95/p99 latency within 2 %") or within four
combined standard errors, whichever is wider.  CPU tier: the engine's state machine on the host
twin; GPU tier: the CUDA engine."""

from __future__ import annotations

import json

import numpy as np
import pytest
from helpers import GOLD, SEED, load_scenario

from asyncflow_b200.flatten import flatten

REF = json.loads((GOLD / "reference_numpy_stats.json").read_text())
N_ENGINE = 600


def compare(name: str, mine: dict[str, np.ndarray]) -> None:
    ref = REF[name]
    for key, vals in mine.items():
        m, sem = float(np.mean(vals)), float(np.std(vals, ddof=1) / np.sqrt(len(vals)))
        r, rsem = ref["mean"][key], ref["sem"][key]
        tol = max(0.02 * abs(r), 4.0 * np.hypot(sem, rsem))
        assert abs(m - r) <= tol, (name, key, m, r, tol)


def summarise(stats, thr, samp_sum, flat) -> dict[str, np.ndarray]:
    c = stats["completed"].astype(np.float64)
    mean = stats["lat_sum"] / c
    ticks = stats["n_ticks"].astype(np.float64)
    return {
        "mean": mean, "median": stats["p50"], "p95": stats["p95"], "p99": stats["p99"],
        "std_dev": np.sqrt(np.maximum(stats["lat_sumsq"] / c - mean * mean, 0.0)),
        "total_requests": c, "generated": stats["generated"].astype(np.float64),
        "rps_mean": thr.sum(axis=1) / flat.horizon_s,
        "ram_mean_first_server": samp_sum[:, 2] / ticks,
        "io_mean_first_server": samp_sum[:, 1] / ticks,
    }


@pytest.mark.parametrize("name", sorted(REF))
def test_twin_matches_reference_statistics(name):
    import twin
    flat = flatten(load_scenario(name, REF[name]["horizon"]))
    r = twin.run(flat, seed=SEED + 1, replica_begin=0, n=N_ENGINE)
    assert (r["stats"]["flags"] == 0).all()
    compare(name, summarise(r["stats"], r["thr"], r["samp_sum"], flat))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(REF))
def test_engine_matches_reference_statistics(name):
    from asyncflow_b200 import SweepRunner
    flat = flatten(load_scenario(name, REF[name]["horizon"]))
    sw = SweepRunner(flat, 4096, seed=SEED + 1, throughput=True)
    res = sw.run()
    sw.close()
    assert not res.overflowed.any()
    compare(name, summarise(res.stats, res.throughput, res.samp_sum, flat))
