"""The reference's own end-to-end acceptance bands, applied to the CUDA engine (and to its CPU twin).

The reference's system tests (tests/system/*.py, enabled with ASYNCFLOW_RUN_SYSTEM_TESTS=1) are the
only end-to-end pins upstream has: loose statistical bands on one unseeded run.  Here the same
scenarios run as 512-replica sweeps, every replica has to sit inside the band, and the Monte-Carlo
mean has to sit close to the analytic value the band was built around.
Scenario parameters are the ones of the reference tests (cited per test)."""

from __future__ import annotations

import copy

import numpy as np
import pytest
from helpers import SEED

from asyncflow_b200 import GpuSimulationRunner, SweepRunner, flatten

N = 512


@pytest.fixture(autouse=True, params=[pytest.param("twin"), pytest.param("gpu", marks=pytest.mark.gpu)])
def engine_kind(request, monkeypatch):
    """Every band is checked twice: on the CUDA engine (GPU box) and, with the identical runner code, on
    the CPU twin standing in for it (tests/twin_engine.py) -- the same state machine, so the statistics the
    bands test are available without a GPU too."""
    if request.param == "twin":
        import asyncflow_b200.runner as R
        from twin_engine import TwinEngine
        monkeypatch.setattr(R, "Engine", TwinEngine)
    return request.param


def _server(sid: str) -> dict:
    return {"id": sid, "server_resources": {"cpu_cores": 1, "ram_mb": 2048},
            "endpoints": [{"endpoint_name": "/api", "steps": [
                {"kind": "initial_parsing", "step_operation": {"cpu_time": 0.001}},
                {"kind": "ram", "step_operation": {"necessary_ram": 64}},
                {"kind": "io_wait", "step_operation": {"io_waiting_time": 0.010}}]}]}


def _exp(mean: float) -> dict:
    return {"mean": mean, "distribution": "exponential"}


def single_server(horizon: int, users: int = 80) -> dict:
    """reference tests/system/test_sys_single_server.py:55-121"""
    return {
        "rqs_input": {"id": "rqs-1", "avg_active_users": {"mean": users},
                      "avg_request_per_minute_per_user": {"mean": 20}, "user_sampling_window": 60},
        "topology_graph": {"nodes": {"client": {"id": "client-1"}, "servers": [_server("srv-1")]},
                           "edges": [{"id": "gen-client", "source": "rqs-1", "target": "client-1", "latency": _exp(0.003)},
                                     {"id": "client-srv", "source": "client-1", "target": "srv-1", "latency": _exp(0.002)},
                                     {"id": "srv-client", "source": "srv-1", "target": "client-1", "latency": _exp(0.003)}]},
        "sim_settings": {"total_simulation_time": horizon, "sample_period_s": 0.05},
    }


def lb_two_servers(horizon: int, users: int = 120) -> dict:
    """reference tests/system/test_sys_lb_two_servers.py:60-150"""
    return {
        "rqs_input": {"id": "rqs-1", "avg_active_users": {"mean": users},
                      "avg_request_per_minute_per_user": {"mean": 20}, "user_sampling_window": 60},
        "topology_graph": {
            "nodes": {"client": {"id": "client-1"},
                      "load_balancer": {"id": "lb-1", "algorithms": "round_robin", "server_covered": ["srv-1", "srv-2"]},
                      "servers": [_server("srv-1"), _server("srv-2")]},
            "edges": [{"id": "gen-client", "source": "rqs-1", "target": "client-1", "latency": _exp(0.003)},
                      {"id": "client-lb", "source": "client-1", "target": "lb-1", "latency": _exp(0.002)},
                      {"id": "lb-srv1", "source": "lb-1", "target": "srv-1", "latency": _exp(0.002)},
                      {"id": "lb-srv2", "source": "lb-1", "target": "srv-2", "latency": _exp(0.002)},
                      {"id": "srv1-client", "source": "srv-1", "target": "client-1", "latency": _exp(0.003)},
                      {"id": "srv2-client", "source": "srv-2", "target": "client-1", "latency": _exp(0.003)}]},
        "sim_settings": {"total_simulation_time": horizon, "sample_period_s": 0.05},
    }


def sweep(payload: dict, n: int = N):
    sw = SweepRunner(flatten(payload), n, seed=SEED, throughput=True)
    res = sw.run()
    sw.close()
    assert not res.overflowed.any()
    return res


def test_single_server_bands():
    """test_sys_single_server.py:124-151: mean latency in [15, 60] ms, mean rps within 35 % of 26.7."""
    res = sweep(single_server(400))
    mean = res.mean_latency
    assert (mean >= 0.015).all() and (mean <= 0.060).all()
    rps = res.throughput.mean(axis=1)
    lam = 80 * 20 / 60.0
    assert (np.abs(rps - lam) / lam <= 0.35).all()
    # Monte-Carlo mean: sum of edge means + 1 ms CPU + 10 ms IO, three hops surviving 1 % dropout each
    assert abs(mean.mean() - (0.003 + 0.002 + 0.003 + 0.011)) < 0.0004
    assert abs(rps.mean() - lam * 0.99 ** 3) / lam < 0.01


def test_lb_two_servers_bands_and_balance():
    """test_sys_lb_two_servers.py:160-204: mean in [20, 60] ms, rps within 30 % of 40, the two LB
    edges' mean concurrency and the two servers' mean RAM within 25 % of each other."""
    payload = lb_two_servers(600)
    flat = flatten(payload)
    res = sweep(payload)
    mean = res.mean_latency
    assert (mean >= 0.020).all() and (mean <= 0.060).all()
    rps = res.throughput.mean(axis=1)
    assert (np.abs(rps - 40.0) / 40.0 <= 0.30).all()
    sm = res.sampled_mean()
    e1, e2 = (3 * flat.n_servers + flat.edge_ids.index(e) for e in ("lb-srv1", "lb-srv2"))
    rel = lambda a, b: np.abs(a - b) / np.maximum(np.maximum(a, b), 1e-12)  # noqa: E731
    assert (rel(sm[:, e1], sm[:, e2]) <= 0.25).all()
    ram1, ram2 = sm[:, 0 * 3 + 2], sm[:, 1 * 3 + 2]
    assert (rel(ram1, ram2) <= 0.25).all()
    # one traced replica answers the analyzer API the reference test uses
    one = GpuSimulationRunner(simulation_input=lb_two_servers(60), seed=SEED).run()
    assert set(one.list_server_ids()) == {"srv-1", "srv-2"}
    sampled = one.get_sampled_metrics()
    assert "lb-srv1" in sampled["edge_concurrent_connection"] and "srv-2" in sampled["ram_in_use"]


def test_edge_spike_raises_mean_latency():
    """test_sys_ev_inj_single_server.py:158-199: +50 ms on client->srv during [0.5, 2.5] s of a 100 s
    run raises the mean by >= 2 % and leaves throughput within 20 %."""
    base = single_server(100)
    spiked = copy.deepcopy(base)
    spiked["events"] = [{"event_id": "net-spike-1", "target_id": "client-srv",
                         "start": {"kind": "network_spike_start", "t_start": 0.5, "spike_s": 0.050},
                         "end": {"kind": "network_spike_end", "t_end": 2.5}}]
    a, b = sweep(base), sweep(spiked)
    # same seed, same replica ids: identical arrivals, so the comparison is paired
    np.testing.assert_array_equal(a.generated, b.generated)
    assert b.mean_latency.mean() >= 1.02 * a.mean_latency.mean()
    ra, rb = a.throughput.mean(axis=1).mean(), b.throughput.mean(axis=1).mean()
    assert abs(rb - ra) / ra <= 0.20
    # 2 s of +50 ms out of 100 s: the mean moves by ~1 ms
    assert abs((b.mean_latency.mean() - a.mean_latency.mean()) - 0.050 * 2.0 / 100.0) < 0.0003


def test_lb_spike_and_outage():
    """test_sys_ev_inj_lb_two_servers.py:190-237: spike on lb->srv-1 [2, 12] s (+50 ms) and srv-2 down
    [5, 20] s in a 100 s run: mean latency >= baseline + 3 ms, throughput in [30 %, 105 %] of baseline."""
    base = lb_two_servers(100, users=80)
    base["topology_graph"]["edges"][2]["id"] = "lb-srv-1"
    base["topology_graph"]["edges"][3]["id"] = "lb-srv-2"
    ev = copy.deepcopy(base)
    ev["events"] = [
        {"event_id": "net-spike", "target_id": "lb-srv-1",
         "start": {"kind": "network_spike_start", "t_start": 2.0, "spike_s": 0.050}, "end": {"kind": "network_spike_end", "t_end": 12.0}},
        {"event_id": "srv2-outage", "target_id": "srv-2",
         "start": {"kind": "server_down", "t_start": 5.0}, "end": {"kind": "server_up", "t_end": 20.0}}]
    a, b = sweep(base), sweep(ev)
    assert b.mean_latency.mean() >= a.mean_latency.mean() + 0.003
    ra, rb = a.throughput.mean(axis=1).mean(), b.throughput.mean(axis=1).mean()
    assert rb / ra >= 0.30 and abs(rb - ra) / ra <= 0.05
    # during the outage every request goes through lb-srv-1: its edge carries more than half
    flat = flatten(ev)
    s1, s2 = flat.edge_ids.index("lb-srv-1"), flat.edge_ids.index("lb-srv-2")
    assert (b.edge_sent[:, s1] > b.edge_sent[:, s2]).all()
    frac = b.edge_sent[:, s2].sum() / (b.edge_sent[:, s1].sum() + b.edge_sent[:, s2].sum())
    assert abs(frac - (100 - 15) / 2 / 100) < 0.01      # srv-2 serves half of the 85 s it is up
