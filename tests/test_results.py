"""Host-side result objects (the OUT side of the boundary), fed with engine-shaped arrays."""

from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import pytest
import twin
from helpers import SEED, load_scenario

from asyncflow_b200 import _capi as K
from asyncflow_b200.flatten import SweepSpec, flatten
from asyncflow_b200.results import ReplicaResults, SweepResults
from asyncflow_b200.runner import arrivals_bound


def _replica(clocks, horizon=2):
    flat = flatten(load_scenario("c1_my_service.yml", horizon))
    flat.horizon_s = horizon
    return ReplicaResults(flat=flat, clocks=np.asarray(clocks, dtype=np.float64).reshape(-1, 2),
                          series=np.zeros((flat.n_series, 0), dtype=np.uint32), generated=0, edge_sent={},
                          edge_dropped={}, n_events=0, flags=0)


def test_throughput_buckets_known_answer():
    # reference tests/unit/metrics/test_analyzer.py:167-175: completions at 1 s and 2 s, window 0.5 -> [0,2,0,2]
    res = _replica([[0.0, 1.0], [0.0, 2.0]])
    ts, rps = res.get_throughput_series(window_s=0.5)
    assert ts == [0.5, 1.0, 1.5, 2.0] and rps == [0.0, 2.0, 0.0, 2.0]
    ts, rps = res.get_throughput_series()
    assert ts == [1.0, 2.0] and rps == [1.0, 1.0]


def test_latency_stats_keys_and_values_follow_the_analyzer():
    res = _replica([[0.0, 0.010], [0.5, 0.530], [1.0, 1.020]])
    st = res.get_latency_stats()
    assert list(st) == ["total_requests", "mean", "median", "std_dev", "p95", "p99", "min", "max"]
    lat = np.array([0.010, 0.030, 0.020])
    assert st["mean"] == float(np.mean(res.latencies)) and abs(st["mean"] - lat.mean()) < 1e-15
    assert st["p95"] == float(np.percentile(res.latencies, 95))
    assert _replica(np.zeros((0, 2))).get_latency_stats() == {}
    assert _replica(np.zeros((0, 2))).format_latency_stats() == "Latency stats: (empty)"
    assert res.format_latency_stats().splitlines()[0].startswith("════════ LATENCY STATS")


def test_sweep_results_summaries():
    flat = flatten(load_scenario("c1_my_service.yml", 6))
    n = 12
    spec = SweepSpec(flat, n, {("users_mean",): np.linspace(20, 300, n)})
    r = twin.run(flat, seed=SEED, n=n, sweep=spec)
    res = SweepResults(flat, r["stats"], r["sent"], r["dropped"], r["samp_sum"], r["samp_max"], r["thr"], r["hist"])
    assert len(res) == n and not res.overflowed.any()
    np.testing.assert_array_equal(res.throughput.sum(axis=1), res.completed)
    i = 7
    st = res.latency_stats(i)
    assert st["total_requests"] == res.completed[i] and st["min"] <= st["median"] <= st["p95"] <= st["p99"] <= st["max"] * 1.02
    assert abs(st["mean"] - r["stats"]["lat_sum"][i] / r["stats"]["completed"][i]) < 1e-15
    sm = res.sampled_mean()
    assert sm.shape == (n, flat.n_series) and (sm[:, 2] <= 2048).all()          # RAM in use never exceeds the server's RAM
    s = res.summary()
    assert s["replicas"] == n and s["completed"] == float(res.completed.sum()) and s["overflowed"] == 0
    est, lo, hi = res.confidence_interval("completed")
    assert lo < est < hi and est == res.completed.mean()
    est2, lo2, hi2 = res.confidence_interval("p95", np.arange(n) < 4)
    assert np.isfinite([est2, lo2, hi2]).all()
    with pytest.raises(KeyError):
        res.confidence_interval("bogus")
    both = SweepResults.concatenate([res, res])
    assert len(both) == 2 * n and both.completed[:n].tolist() == res.completed.tolist()
    # more users -> more completions (the server is not saturated at 300 users x 100 rpm for 6 s? it is: just monotone in the low range)
    assert res.completed[0] < res.completed[3] < res.completed[6]


def test_overflow_flags_surface_in_sweep_results():
    flat = flatten(load_scenario("overload_single.yml"))
    r = twin.run(flat, seed=SEED, n=2, request_capacity=150)
    res = SweepResults(flat, r["stats"], r["sent"], r["dropped"], r["samp_sum"], r["samp_max"])
    assert res.overflowed.all() and res.summary()["overflowed"] == 2
    assert (r["stats"]["flags"] & K.FLAG_REQUEST_OVERFLOW).all()


def test_arrivals_bound_is_generous():
    flat = flatten(load_scenario("c1_my_service.yml", 60))
    b = arrivals_bound(flat)
    assert b > 100 * (100 / 60) * 60 * 1.5          # mean arrivals ~10 000: the bound leaves > 50 % head-room
    assert arrivals_bound(flat, users_mean=1000.0) > 1000 * (100 / 60) * 60


def test_series_bands_across_traced_replicas_match_numpy():
    """per-tick mean/min/max of a sampled series over the traced replicas of a sweep."""
    from asyncflow_b200 import series_bands
    from asyncflow_b200.flatten import SweepSpec

    payload = load_scenario("c3_lb_two_servers.yml", 12)
    flat = flatten(payload)
    k = 6
    spec = SweepSpec(flat, k, {("users_mean",): np.linspace(100, 600, k)})
    r = twin.run(flat, seed=SEED, replica_begin=0, n=k, sweep=spec, trace=k, clock_cap=100000)
    reps = []
    for i in range(k):
        st = r["stats"][i]
        n, nt = int(st["completed"]), int(st["n_ticks"])
        reps.append(ReplicaResults(flat=flat, clocks=r["trace_clocks"][i, :n].copy(), series=r["trace_series"][i][:, :nt].copy(),
                                   generated=int(st["generated"]), edge_sent={}, edge_dropped={}, n_events=int(st["n_events"]),
                                   flags=int(st["flags"])))
    b = series_bands(reps, "ram_in_use", "srv-1")
    raw = np.stack([r["trace_series"][i][3 * flat.server_ids.index("srv-1") + 2, :len(b["t"])] for i in range(k)]).astype(np.float64)
    assert b["n"] == k and len(b["t"]) == int(r["stats"][0]["n_ticks"])
    np.testing.assert_array_equal(b["mean"], raw.mean(axis=0))
    np.testing.assert_array_equal(b["min"], raw.min(axis=0))
    np.testing.assert_array_equal(b["max"], raw.max(axis=0))
    assert b["max"].max() > b["min"].max()                         # the sweep does spread the curves
    np.testing.assert_allclose(b["t"][:3], [0.0, flat.sample_period, 2 * flat.sample_period])
    # edge series, enum-like keys and unknown entities behave like ResultsAnalyzer.get_series
    e = series_bands(reps, SimpleNamespace(value="edge_concurrent_connection"), flat.edge_ids[0])
    assert e["n"] == k and e["max"].max() >= 1
    assert series_bands(reps, "ram_in_use", "nope")["n"] == 0
    # a SweepResults carries them along
    sw = SweepResults(flat, r["stats"], r["sent"], r["dropped"], r["samp_sum"], r["samp_max"])
    sw.traced = reps
    np.testing.assert_array_equal(sw.take(np.arange(k)[::-1]).bands("ram_in_use", "srv-1")["mean"], b["mean"])
