"""SweepRunner's host-side orchestration, end to end on the CPU twin (tests/twin_engine.py stands in for
the ctypes engine): plain runs, shard ranges, heaviest-first launch order, traced replicas, drill-down."""

from __future__ import annotations

import des_port
import numpy as np
from helpers import SEED, load_scenario
from twin_engine import TwinEngine

from asyncflow_b200 import SweepRunner
from asyncflow_b200.distributed import shard_bounds

USERS = [40.0, 250.0, 15.0, 120.0, 70.0, 180.0, 30.0]
RTT = [0.001, 0.002, 0.003, 0.004, 0.005, 0.006, 0.007]


def runner(**kw) -> SweepRunner:
    base = load_scenario("c1_my_service.yml", 6)
    sw = SweepRunner(base, len(USERS), {("users_mean",): USERS, ("edge_mean", "client-app"): RTT}, seed=SEED,
                     pinned=False, throughput=True, **kw)
    sw._engine = TwinEngine()
    sw._engine.upload(sw.flat)
    return sw


def oracle_row(sw: SweepRunner, row: int) -> dict:
    return des_port.simulate(sw.payload_for(row), seed=SEED, replica=int(sw.replica_ids[row]))


def test_plain_run_returns_one_row_per_sweep_row():
    sw = runner()
    res = sw.run()
    assert len(res) == len(USERS) and res.rows.tolist() == list(range(len(USERS)))
    for row in range(len(USERS)):
        o = oracle_row(sw, row)
        assert int(res.generated[row]) == o["generated"] and int(res.completed[row]) == o["completed"]
        assert dict(zip(sw.flat.edge_ids, map(int, res.edge_dropped[row]))) == o["edge_dropped"]
    assert sw.h2d_bytes == len(USERS) * 2 * 8 and sw.d2h_bytes > 0


def test_balanced_run_comes_back_in_row_order_with_the_ids_it_names():
    sw = runner(balance=True)
    assert sw.order.tolist() == [1, 5, 3, 4, 0, 6, 2]
    res = sw.run()
    assert res.rows.tolist() == list(range(len(USERS)))
    gen = res.generated.astype(np.int64)
    assert np.argmax(gen) == 1 and np.argmin(gen) == 2           # users 250 / users 15: rows, not launch positions
    for row in range(len(USERS)):
        o = oracle_row(sw, row)                                   # replica id = replica_ids[row]
        assert int(res.generated[row]) == o["generated"] and int(res.completed[row]) == o["completed"]
        assert abs(float(res.stats["lat_sum"][row]) - sum(b - a for a, b in o["clocks"])) < 1e-9
    # the per-second throughput rows were un-permuted with everything else
    assert int(res.throughput[1].sum()) == int(res.completed[1])


def test_shards_of_a_dealt_sweep_cover_every_row_once():
    world = 2
    sw_all = runner(balance=True, deal=world)
    full = sw_all.run()
    seen = []
    for rank in range(world):
        sw = runner(balance=True, deal=world)
        b, e = shard_bounds(len(USERS), rank, world)
        part = sw.run(b, e)                                       # a shard: id order, .rows names the sweep rows
        assert part.rows.tolist() == sw.order[b:e].tolist()
        for j, row in enumerate(part.rows):
            assert int(part.generated[j]) == int(full.generated[row])     # independent of the world size
            assert float(part.stats["lat_sum"][j]) == float(full.stats["lat_sum"][row])
        seen += part.rows.tolist()
        cost = np.array(USERS)[part.rows]
        assert (np.diff(cost) <= 0).all()                         # heaviest first inside the shard
    assert sorted(seen) == list(range(len(USERS)))
    calls = [c for c in sw._engine.calls if c[0] == "upload_sweep"]
    assert calls[-1][1:] == (b, b, e - b)


def test_traced_replicas_and_bands():
    sw = runner(trace_replicas=3)
    res = sw.run()
    cfg = [c for c in sw._engine.calls if c[0] == "configure"][-1][1]
    assert cfg["trace_replicas"] == 3 and cfg["trace_clock_capacity"] == sw.request_capacity
    assert len(res.traced) == 3
    for j, rep in enumerate(res.traced):
        o = oracle_row(sw, j)
        assert rep.clocks.shape == (o["completed"], 2)
        assert [tuple(x) for x in rep.clocks.tolist()] == [tuple(c) for c in o["clocks"]]
        assert rep.get_latency_stats()["total_requests"] == o["completed"]
    band = res.bands("ram_in_use", "app-1")
    assert band["n"] == 3 and (band["min"] <= band["mean"]).all() and (band["mean"] <= band["max"]).all()
    # drill-down replays the same replica
    rr = sw.replica_runner(1)
    assert (rr.seed, rr.replica) == (SEED, 1) and rr.simulation_input["rqs_input"]["avg_active_users"]["mean"] == USERS[1]
    # an untraced sweep configures no tracing at all (the bench path)
    plain = runner()
    plain.run()
    cfg = [c for c in plain._engine.calls if c[0] == "configure"][-1][1]
    assert cfg["trace_replicas"] == 0 and cfg["trace_clock_capacity"] == 0


def test_single_replica_runner_on_the_twin_engine(monkeypatch):
    """GpuSimulationRunner.run(): capacity retry loop and the ReplicaResults it assembles."""
    import asyncflow_b200.runner as R
    engines = []

    def make(device=0):
        e = TwinEngine(device)
        engines.append(e)
        return e
    monkeypatch.setattr(R, "Engine", make)
    payload = load_scenario("overload_single.yml")
    monkeypatch.setattr(R, "arrivals_bound", lambda flat, *a: 300)      # far too small: forces the x4 retries
    res = R.GpuSimulationRunner(simulation_input=payload, seed=SEED, replica=5).run()
    o = des_port.simulate(payload, seed=SEED, replica=5)
    assert [tuple(x) for x in res.clocks.tolist()] == [tuple(c) for c in o["clocks"]]
    assert res.generated == o["generated"] and res.edge_dropped == o["edge_dropped"] and res.flags == 0
    runs = [c for c in engines[0].calls if c[0] == "run"]
    caps = [c[1]["request_capacity"] for c in engines[0].calls if c[0] == "configure"]
    assert len(runs) >= 2 and caps == sorted(caps) and caps[-1] >= 4 * caps[0]
    import pytest
    with pytest.raises(RuntimeError):
        r = R.GpuSimulationRunner(simulation_input=payload, seed=SEED)
        r._ran = True
        r.run()


def test_overflowed_replicas_are_rerun_with_larger_pools():
    """Tiny pools: the saturated rows overflow (flagged); retry_overflow re-runs just them, exactly."""
    base = load_scenario("c1_my_service.yml", 6)
    users = [30.0, 900.0, 40.0, 800.0, 850.0, 20.0]

    def make(**kw):
        sw = SweepRunner(base, len(users), {("users_mean",): users}, seed=SEED, pinned=False, throughput=True,
                         request_capacity=300, event_capacity=64, **kw)
        sw._engine = TwinEngine()
        sw._engine.upload(sw.flat)
        return sw
    plain = make().run()
    assert plain.overflowed.tolist() == [False, True, False, True, True, False]
    sw = make()
    res = sw.run(retry_overflow=True)
    assert not res.overflowed.any()
    ranges = [(c[2], c[3]) for c in sw._engine.calls if c[0] == "run"]
    assert ranges[0] == (0, 6) and set(ranges[1:]) >= {(1, 2), (3, 5)}      # only the flagged ids, grouped
    for row in range(len(users)):
        o = oracle_row(sw, row)
        assert int(res.generated[row]) == o["generated"] and int(res.completed[row]) == o["completed"]
        assert int(res.throughput[row].sum()) == o["completed"]
    # the rows that never overflowed were not touched
    for row in (0, 2, 5):
        assert plain.stats[row].tobytes() == res.stats[row].tobytes()
    # with a balanced launch order the patched rows still come back in row order
    bal = make(balance=True)
    rb = bal.run(retry_overflow=True)
    assert not rb.overflowed.any() and np.argmax(rb.generated) == 1


def test_frame_lists_parameters_next_to_statistics():
    pd = __import__("pytest").importorskip("pandas")
    sw = runner(balance=True)
    res = sw.run()
    df = sw.frame(res)
    assert list(df["row"]) == list(range(len(USERS)))
    assert list(df["users_mean"]) == USERS and list(df["edge_mean:client-app"]) == RTT
    assert list(df["replica_id"]) == sw.replica_ids.tolist()
    assert (df["completed"] <= df["generated"]).all() and df["p95_s"].notna().all()
    part = sw.run(2, 5)                                   # a shard: rows are named, parameters follow them
    dfp = sw.frame(part)
    assert list(dfp["row"]) == sw.order[2:5].tolist()
    assert list(dfp["users_mean"]) == [USERS[r] for r in sw.order[2:5]]
    assert isinstance(df, pd.DataFrame)


def test_from_yaml_mirrors_the_reference_constructor(monkeypatch, tmp_path):
    """GpuSimulationRunner.from_yaml(env=, yaml_path=) -- the reference's convenience constructor
    (runtime/simulation_runner.py:381-398): reads the YAML, builds the runner, run() gives analyzer-shaped
    results equal to the oracle's on the same payload."""
    import asyncflow_b200.runner as R
    import yaml
    monkeypatch.setattr(R, "Engine", lambda device=0: TwinEngine(device))
    payload = load_scenario("c1_my_service.yml", 12)
    path = tmp_path / "scenario.yml"
    path.write_text(yaml.safe_dump(payload))
    env = object()                                      # accepted and kept, like simpy.Environment there
    runner = R.GpuSimulationRunner.from_yaml(env=env, yaml_path=path, seed=SEED, replica=3)
    assert runner.env is env
    res = runner.run()
    o = des_port.simulate(payload, seed=SEED, replica=3)
    assert [tuple(x) for x in res.clocks.tolist()] == [tuple(c) for c in o["clocks"]]
    lat = res.get_latency_stats()
    assert lat["total_requests"] == o["completed"]
    assert str(path) and R.GpuSimulationRunner.from_yaml(yaml_path=str(path)).simulation_input is not None   # str paths too


def test_results_of_two_shards_do_not_share_memory():
    """run(begin, end) per shard, then concatenate: the second collect() must not overwrite the first
    (the staging buffers are reused; results own their arrays)."""
    from asyncflow_b200.results import SweepResults
    sw = runner()
    a = sw.run(0, 2)
    before = a.completed.copy()
    b = sw.run(2, 4)
    assert not np.shares_memory(a.stats, b.stats) and not np.shares_memory(a.edge_sent, b.edge_sent)
    assert a.completed.tolist() == before.tolist()
    both = SweepResults.concatenate([a, b])
    assert both.completed.tolist() == before.tolist() + b.completed.tolist()


def test_exact_percentiles_replace_the_histogram_ones_for_the_rows_asked_for():
    """SweepRunner.exact_percentiles(res, rows): the rows are simulated again with their (start, finish) lists kept and
    numpy.percentile applied -- the reference analyzer's arithmetic (metrics/analyzer.py:93-103) -- also under
    balance=True, where a row's replica id is not its row number."""
    for kw in ({}, {"balance": True}):
        sw = runner(**kw)
        res = sw.run()
        coarse = res.stats[["p50", "p95", "p99"]].copy()
        rows = [1, 2, 5]
        sw.exact_percentiles(res, rows, chunk=2)
        assert res.exact.tolist() == [i in rows for i in range(len(USERS))]
        for row in range(len(USERS)):
            o = oracle_row(sw, row)
            lat = np.array([b - a for a, b in o["clocks"]])
            for q, key in ((50, "p50"), (95, "p95"), (99, "p99")):
                exact = float(np.percentile(lat, q))
                if row in rows:
                    assert float(res.stats[key][row]) == exact, (kw, row, key)
                else:
                    assert float(res.stats[key][row]) == float(coarse[key][row])
                    assert abs(float(res.stats[key][row]) - exact) <= 0.01 * exact

