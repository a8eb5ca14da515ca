"""CPU tier: the thread-per-replica state machine (asyncflow_b200/csrc/af_lane.cuh, compiled for the host as a warp of
one lane) against the oracle, against the warp-per-replica state machine (af_core.cuh), and across its memory tiers."""

from __future__ import annotations

import des_port
import fuzz
import numpy as np
import pytest
import twin
from helpers import PARITY_CASES, SEED, assert_matches_oracle, load_scenario

from asyncflow_b200 import _capi as K
from asyncflow_b200 import SweepSpec, flatten

FIELDS = ("n_events", "generated", "completed", "flags", "n_ticks", "peak_events", "peak_requests",
          "lat_sum", "lat_sumsq", "lat_min", "lat_max", "p50", "p95", "p99")


def _same(a: dict, b: dict) -> None:
    for f in FIELDS:
        np.testing.assert_array_equal(a["stats"][f], b["stats"][f], err_msg=f)
    for k in ("sent", "dropped", "hist", "thr", "samp_sum", "samp_max", "trace_counts"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


@pytest.mark.parametrize("name", sorted(PARITY_CASES))
def test_lane_and_warp_state_machines_agree_on_everything(name):
    """Same stats block (incl. peaks and event counts), counters, histograms, buckets, sampled aggregates and traces."""
    flat = flatten(load_scenario(name, PARITY_CASES[name]))
    kw = dict(seed=SEED, replica_begin=40, n=3, trace=3, clock_cap=300000, request_capacity=400000, event_capacity=8192)
    lane = twin.run(flat, engine="lane", **kw)
    warp = twin.run(flat, engine="warp", **kw)
    _same(lane, warp)
    for j in range(3):
        n, nt = int(lane["stats"][j]["completed"]), int(lane["stats"][j]["n_ticks"])
        np.testing.assert_array_equal(lane["trace_clocks"][j, :n], warp["trace_clocks"][j, :n])
        np.testing.assert_array_equal(lane["trace_series"][j][:, :nt], warp["trace_series"][j][:, :nt])


@pytest.mark.parametrize("lane_bytes", [0, 300, 908, 1816, 7000])
def test_results_do_not_depend_on_the_shared_memory_tier_split(lane_bytes):
    """lane_bytes = the lane's share of shared memory (0: the smallest the scenario admits): however the events,
    requests and now-queue items are split between the two tiers, the replica is the same."""
    for name in ("c3_lb_two_servers.yml", "overload_single.yml", "poisson_ties.yml", "c5_multihop32.yml"):
        payload = load_scenario(name, PARITY_CASES[name])
        flat = flatten(payload)
        r = twin.run(flat, engine="lane", lane_bytes=lane_bytes or 1, seed=SEED, replica_begin=11, n=1, trace=1, clock_cap=300000)
        o = des_port.simulate(payload, seed=SEED, replica=11)
        n, nt = int(r["stats"][0]["completed"]), int(r["stats"][0]["n_ticks"])
        assert r["stats"][0]["flags"] == 0
        assert_matches_oracle(o, flat, stats=r["stats"][0], clocks=r["trace_clocks"][0, :n], sent=r["sent"][0],
                              dropped=r["dropped"][0], series=r["trace_series"][0][:, :nt], throughput=r["thr"][0],
                              hist=r["hist"][0])


@pytest.mark.parametrize("name,lane_bytes", [("c3_lb_two_servers.yml", 648), ("c4_lb8_events.yml", 1036),
                                             ("c5_multihop32.yml", 1444), ("overload_single.yml", 660)])
def test_results_do_not_depend_on_the_split_between_events_and_records(name, lane_bytes):
    """ev_need = the pending-events estimate af_run splits a lane's shared memory by (0: evenly): events first, records
    down to 2 slots -- at the budgets the CUDA engine really runs (600-1450 B per lane) the replica is the same."""
    payload = load_scenario(name, PARITY_CASES[name])
    flat = flatten(payload)
    o = des_port.simulate(payload, seed=SEED, replica=23)
    for ev_need in (0, 4, 26, 60, 100000):
        r = twin.run(flat, engine="lane", lane_bytes=lane_bytes, ev_need=ev_need, seed=SEED, replica_begin=23, n=1, trace=1,
                     clock_cap=300000)
        n, nt = int(r["stats"][0]["completed"]), int(r["stats"][0]["n_ticks"])
        assert r["stats"][0]["flags"] == 0, ev_need
        assert_matches_oracle(o, flat, stats=r["stats"][0], clocks=r["trace_clocks"][0, :n], sent=r["sent"][0],
                              dropped=r["dropped"][0], series=r["trace_series"][0][:, :nt], throughput=r["thr"][0],
                              hist=r["hist"][0])


def test_random_sweeps_on_the_lane_engine_match_the_oracle_row_by_row():
    """Every AF_FIELD_* of the C ABI as a sweep column (fields read at the start of a replica and fields looked up
    through the lane's row copy), tiny shared-memory tier."""
    for seed in range(900, 912):
        payload = fuzz.scenario(seed)
        flat = flatten(payload)
        n = 3
        spec = SweepSpec(flat, n, fuzz.sweep_columns(seed, payload, n))
        r = twin.run(flat, engine="lane", lane_bytes=400, seed=SEED, n=n, sweep=spec, trace=n, clock_cap=100000,
                     request_capacity=200000)
        for i in range(n):
            p = spec.payload_for(payload, i)
            o = des_port.simulate(p, seed=SEED, replica=i)
            m = int(r["stats"][i]["completed"])
            assert r["stats"][i]["flags"] == 0
            assert_matches_oracle(o, flatten(p), stats=r["stats"][i], clocks=r["trace_clocks"][i, :m], sent=r["sent"][i],
                                  dropped=r["dropped"][i])


def test_columns_that_repeat_each_other_share_a_row_slot_and_change_nothing():
    """configs[2]'s shape: one RTT array and one jitter array over every edge (12 columns, 2 distinct value arrays).
    The lane engine keeps one row slot per distinct array (aflh::column_aliases); rows still equal the oracle on the
    row's own payload -- and a run that reaches past the sweep's rows (base values) does not share slots."""
    import bench
    payload = bench.workload(0, 6)
    flat = flatten(payload)
    n = 4
    rtt = np.array([0.001, 0.013, 0.027, 0.05]); sig = np.array([0.1, 0.2, 0.35, 0.5]) * rtt
    cols = {}
    for e in flat.edge_ids:
        cols[("edge_mean", e)] = rtt
        cols[("edge_sigma", e)] = sig
    spec = SweepSpec(flat, n, cols)
    r = twin.run(flat, engine="lane", lane_bytes=500, seed=SEED, n=n + 1, sweep=spec, trace=n + 1, clock_cap=100000)
    for i in range(n + 1):
        p = spec.payload_for(payload, i) if i < n else payload
        o = des_port.simulate(p, seed=SEED, replica=i)
        m = int(r["stats"][i]["completed"])
        assert_matches_oracle(o, flatten(p), stats=r["stats"][i], clocks=r["trace_clocks"][i, :m], sent=r["sent"][i],
                              dropped=r["dropped"][i])
    r2 = twin.run(flat, engine="lane", lane_bytes=500, seed=SEED, n=n, sweep=spec, trace=n, clock_cap=100000)
    for i in range(n):
        m = int(r2["stats"][i]["completed"])
        np.testing.assert_array_equal(r2["trace_clocks"][i, :m], r["trace_clocks"][i, :m])


def test_lane_pool_overflow_is_flagged():
    flat = flatten(load_scenario("overload_single.yml"))
    r = twin.run(flat, engine="lane", seed=SEED, n=1, request_capacity=200)
    assert r["stats"][0]["flags"] & K.FLAG_REQUEST_OVERFLOW
    r = twin.run(flat, engine="lane", seed=SEED, n=1, event_capacity=4)
    assert r["stats"][0]["flags"] & K.FLAG_EVENT_OVERFLOW


def _lb_empty_payload(horizon: int = 6):
    """The LB covers only srv-1, srv-1 chains to srv-2; an outage takes srv-1 down.  The reference's validator accepts
    it (srv-2 is still up) and its load balancer then raises on an empty edge set."""
    p = load_scenario("c3_lb_two_servers.yml", horizon)
    topo = p["topology_graph"]
    topo["nodes"]["load_balancer"]["server_covered"] = ["srv-1"]
    topo["edges"] = [e for e in topo["edges"] if e["id"] not in ("lb-srv2", "srv1-client")]
    topo["edges"].append({"id": "srv1-srv2", "source": "srv-1", "target": "srv-2",
                          "latency": {"mean": 0.001, "distribution": "exponential"}})
    p["events"] = [{"event_id": "down-1", "target_id": "srv-1", "start": {"kind": "server_down", "t_start": 2.0},
                    "end": {"kind": "server_up", "t_end": 4.0}}]
    return p


def test_an_outage_that_empties_the_lb_pool_is_rejected_up_front_and_flagged_by_the_engines():
    with pytest.raises(ValueError, match="every server behind the load balancer is down"):
        flatten(_lb_empty_payload())
    # the engines themselves stop the replica instead of reading lb[-1]: a horizon that ends before the outage
    # passes the host check; then the POD's horizon is put back
    flat = flatten(_lb_empty_payload(1))
    flat.pod.horizon_s = 6
    flat.horizon_s = 6
    for engine in ("lane", "warp"):
        r = twin.run(flat, engine=engine, seed=SEED, n=2)
        assert (r["stats"]["flags"] & K.FLAG_LB_EMPTY).all(), engine
        assert (r["stats"]["completed"] > 0).all()


def test_shared_memory_split_follows_the_estimated_need_but_never_starves_a_table():
    """af_run splits a lane's dynamic shared memory between pending events and request records by the number of events a
    replica of the launch typically holds (aflh::pending_events_estimate: Little's law on the scenario and the sweep's
    maxima).  Properties: never fewer events than the even split, never fewer than 2 records, and the
    bench workload (RTT up to 50 ms: ~30 requests in flight) does get more events than the even split."""
    import ctypes as C

    import bench
    L = twin.lib()

    def split(w, budget, with_sweep=True):
        flat = flatten(w.payload)
        sw_p, keep = None, None
        if with_sweep:
            ids = np.arange(0, 2000, dtype=np.int64) * (w.replicas // 2000)
            spec = SweepSpec(flat, len(ids), w.columns(flat, ids, w.replicas))
            sw, keep = spec.pod(0, None)
            sw_p = C.byref(sw)
        out = (C.c_int32 * 2)()
        assert L.af_twin_lane_split(C.byref(flat.pod), sw_p, budget, 0, out) == 0
        even = tuple(out)
        assert L.af_twin_lane_split(C.byref(flat.pod), sw_p, budget, 1, out) == 0
        need = tuple(out)
        est = L.af_twin_pending_events_estimate(C.byref(flat.pod), sw_p)
        del keep
        return est, even, need

    for key, budget in (("c3", 660), ("c3", 904), ("c2", 660), ("c4", 1036), ("c4", 1452), ("c5", 1204)):
        est, even, need = split(bench.make_workload(key), budget)
        assert need[0] >= even[0] and need[1] >= min(2, even[1]) or need == even, (key, budget, est, even, need)
        assert 16 * need[0] + 20 * need[1] <= 16 * even[0] + 20 * even[1] + 36, (key, budget, even, need)
    est, even, need = split(bench.make_workload("c3"), 660)
    assert est >= 25 and need[0] > even[0], (est, even, need)
    est, even, need = split(bench.make_workload("c2"), 1036, with_sweep=False)      # README my_service.yml: ~3 in flight
    assert est <= 12 and need == even, (est, even, need)

