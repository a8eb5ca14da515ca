"""GPU tier: the CUDA engine, called through the C ABI, against the oracle and the
golden vectors (bit-exact), plus size-independent properties at BASELINE sizes."""

from __future__ import annotations

import des_port
import numpy as np
import pytest
from helpers import (PARITY_CASES, SEED, assert_matches_oracle, check_against_golden, load_golden,
                     load_scenario)

from asyncflow_b200 import _capi as K
from asyncflow_b200 import Engine, GpuSimulationRunner, SweepRunner, SweepSpec, flatten

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    with Engine(0) as e:
        yield e


#: pass structures of af_run (af_engine_set_mode): the product default, and each kernel pinned.  The pinned lane
#: kernel takes the capacities as given (no escalation), so it gets the warp engine's defaults.
MODES = {"auto": {}, "two_pass": {}, "lane": {"event_capacity": 2048, "request_capacity": 16384}, "warp": {}}


@pytest.fixture(params=sorted(MODES))
def mode(request, eng):
    eng.set_mode(request.param)
    yield request.param
    eng.set_mode("auto")


def run_traced(eng, flat, begin, n, clock_cap=200000, **kw):
    eng.upload(flat)
    eng.configure(trace_replicas=n, trace_clock_capacity=clock_cap, **kw)
    eng.run(SEED, begin, begin + n)
    st = eng.stats()
    sent, dropped = eng.edge_counts()
    return st, sent, dropped


@pytest.mark.parametrize("name", sorted(PARITY_CASES))
def test_engine_reproduces_golden_vectors(eng, mode, name):
    gold = load_golden(name)
    flat = flatten(load_scenario(name, gold["horizon"]))
    for vec in gold["vectors"]:
        st, sent, dropped = run_traced(eng, flat, vec["replica"], 1, **MODES[mode])
        assert st[0]["flags"] == 0
        passes = eng.last_run_passes()
        if mode != "auto":      # (auto: a handful of replicas goes to the warp-per-replica kernel alone)
            assert passes["lane_pass"] == (mode != "warp") and passes["warp_pass"] == (mode != "lane")
        check_against_golden(
            vec, generated=int(st[0]["generated"]), completed=int(st[0]["completed"]),
            clocks=eng.trace_clocks(0), edge_sent=dict(zip(flat.edge_ids, map(int, sent[0]))),
            edge_dropped=dict(zip(flat.edge_ids, map(int, dropped[0]))),
            throughput=eng.throughput()[0], series=eng.trace_series(0), flat=flat)


@pytest.mark.parametrize("name", ["c1_my_service.yml", "c3_lb_two_servers.yml"])
def test_engine_reproduces_the_baseline_horizons(eng, mode, name):
    """The BASELINE shapes at their BASELINE horizons (60 s / 600 s) against the unmodified reference actors."""
    gold = load_golden(name.replace(".yml", "_full.yml"))
    flat = flatten(load_scenario(name, gold["horizon"]))
    for vec in gold["vectors"]:
        st, sent, dropped = run_traced(eng, flat, vec["replica"], 1, **MODES[mode])
        assert st[0]["flags"] == 0
        check_against_golden(
            vec, generated=int(st[0]["generated"]), completed=int(st[0]["completed"]),
            clocks=eng.trace_clocks(0), edge_sent=dict(zip(flat.edge_ids, map(int, sent[0]))),
            edge_dropped=dict(zip(flat.edge_ids, map(int, dropped[0]))),
            throughput=eng.throughput()[0], series=eng.trace_series(0), flat=flat)


@pytest.mark.parametrize("name", sorted(PARITY_CASES))
def test_engine_matches_oracle_on_fresh_replicas(eng, mode, name):
    horizon = {"c1_my_service.yml": 10, "c3_lb_two_servers.yml": 12, "c4_lb8_events.yml": 245,
               "c5_multihop32.yml": 6}.get(name)
    payload = load_scenario(name, horizon)
    flat = flatten(payload)
    reps = [21, 22, 23] if not name.startswith("c4") else [21]
    st, sent, dropped = run_traced(eng, flat, reps[0], len(reps), **MODES[mode])
    thr, hist = eng.throughput(), eng.histograms()
    ssum, smax = eng.sampled()
    for i, rep in enumerate(reps):
        o = des_port.simulate(payload, seed=SEED, replica=rep)
        series = eng.trace_series(i)
        assert_matches_oracle(o, flat, stats=st[i], clocks=eng.trace_clocks(i), sent=sent[i],
                              dropped=dropped[i], series=series, throughput=thr[i], hist=hist[i])
        np.testing.assert_array_equal(ssum[i], series.astype(np.uint64).sum(axis=1))
        np.testing.assert_array_equal(smax[i], series.max(axis=1) if series.shape[1] else 0)


def test_results_do_not_depend_on_batching_or_launch_shape(eng):
    """Replica r is a function of (seed, r) only: one at a time == inside a batch of 300,
    whatever the warps-per-block / occupancy."""
    flat = flatten(load_scenario("mixed_lc.yml", 12))
    eng.upload(flat)
    eng.configure()
    eng.run(SEED, 100, 400)
    a = eng.stats().copy()
    a_sent, a_drop = eng.edge_counts()
    for wpb, bps in ((1, 1), (8, 0), (3, 2), (2, 5)):
        eng.configure(warps_per_block=wpb, blocks_per_sm=bps)
        eng.run(SEED, 100, 400)
        b = eng.stats()
        for f in ("generated", "completed", "n_events", "lat_sum", "lat_sumsq", "lat_min", "lat_max", "n_ticks", "p95"):
            np.testing.assert_array_equal(a[f], b[f], err_msg=f)
        np.testing.assert_array_equal(a_sent, eng.edge_counts()[0])
    eng.configure()
    eng.run(SEED, 250, 251)
    one = eng.stats()[0]
    for f in ("generated", "completed", "n_events", "lat_sum", "lat_min", "lat_max"):
        assert one[f] == a[150][f]


def test_sweep_rows_match_per_replica_oracle(eng):
    base = load_scenario("c1_my_service.yml", 8)
    flat = flatten(base)
    users = [15.0, 90.0, 300.0, 700.0]
    rtt = [0.001, 0.004, 0.02, 0.05]
    spec = SweepSpec(flat, 4, {("users_mean",): users, ("edge_mean", "client-app"): rtt,
                               ("server_ram_mb", "app-1"): [2048, 256, 512, 1024]})
    eng.upload(flat)
    eng.configure(trace_replicas=4, trace_clock_capacity=60000, request_capacity=60000)
    eng.upload_sweep(spec, 0)
    eng.run(SEED, 0, 4)
    st = eng.stats()
    sent, dropped = eng.edge_counts()
    for i in range(4):
        p = load_scenario("c1_my_service.yml", 8)
        p["rqs_input"]["avg_active_users"]["mean"] = users[i]
        p["topology_graph"]["edges"][1]["latency"]["mean"] = rtt[i]
        p["topology_graph"]["nodes"]["servers"][0]["server_resources"]["ram_mb"] = [2048, 256, 512, 1024][i]
        o = des_port.simulate(p, seed=SEED, replica=i)
        assert_matches_oracle(o, flat, stats=st[i], clocks=eng.trace_clocks(i), sent=sent[i], dropped=dropped[i])
    eng.upload_sweep(None)


def test_overflow_is_reported_per_replica(eng):
    flat = flatten(load_scenario("overload_single.yml"))
    eng.upload(flat)
    eng.configure(request_capacity=300)
    eng.run(SEED, 0, 3)
    assert (eng.stats()["flags"] & K.FLAG_REQUEST_OVERFLOW).all()
    eng.configure()
    eng.run(SEED, 0, 3)
    assert (eng.stats()["flags"] == 0).all()


def test_runner_api_mirrors_the_reference(eng, tmp_path):
    payload = load_scenario("c1_my_service.yml", 10)
    res = GpuSimulationRunner(env=None, simulation_input=payload, seed=SEED, replica=0).run()
    o = des_port.simulate(payload, seed=SEED, replica=0)
    np.testing.assert_array_equal(res.clocks, np.array(o["clocks"]).reshape(-1, 2))
    st = res.get_latency_stats()
    lat = res.latencies
    assert st["total_requests"] == o["completed"] and st["p95"] == float(np.percentile(lat, 95))
    ts, rps = res.get_throughput_series()
    assert len(ts) == 10 and sum(rps) == o["completed"]
    times, vals = res.get_series("ram_in_use", "app-1")
    assert vals == o["server_series"]["app-1"]["ram_in_use"] and times[1] == 0.05
    r2 = GpuSimulationRunner(simulation_input=payload)
    r2.run()
    with pytest.raises(RuntimeError):
        r2.run()


# ---- BASELINE-size properties (no oracle at this size) ----------------------------
def test_c2_sweep_properties_at_full_size():
    """configs[1]: 10 000 replicas of the README topology sweeping users 10..1000."""
    n = 10_000
    flat = flatten(load_scenario("c1_my_service.yml", 20))
    users = np.linspace(10, 1000, n)
    sw = SweepRunner(flat, n, {("users_mean",): users}, seed=SEED, throughput=True)
    res = sw.run()
    st = res.stats
    assert not res.overflowed.any()
    sent, dropped = res.edge_sent, res.edge_dropped
    ge, ca, ac = (flat.edge_ids.index(e) for e in ("gen-client", "client-app", "app-client"))
    # conservation: what the generator made went onto the first edge; every survivor of an edge
    # is either still in flight at the horizon or went onto the next one
    np.testing.assert_array_equal(sent[:, ge], st["generated"])
    assert (sent[:, ca] <= sent[:, ge] - dropped[:, ge]).all()
    assert (sent[:, ac] <= sent[:, ca] - dropped[:, ca]).all()
    assert (st["completed"] <= sent[:, ac] - dropped[:, ac]).all()
    np.testing.assert_array_equal(res.throughput.sum(axis=1), st["completed"])
    # dropout is ~1 % per edge overall
    assert abs(dropped[:, ge].sum() / sent[:, ge].sum() - 0.01) < 0.001
    # arrivals scale with users (rate 100 rpm): generated ~ users * 100/60 * T
    lo, hi = st["generated"][:500].mean(), st["generated"][-500:].mean()
    assert 0.8 < lo / (users[:500].mean() * 100 / 60 * 20) < 1.2
    assert 0.8 < hi / (users[-500:].mean() * 100 / 60 * 20) < 1.2
    # the single core saturates at 500 req/s: overloaded replicas complete < 500 * T
    assert st["completed"].max() <= 500 * 20
    assert np.nanmedian(res.mean_latency[-500:]) > 10 * np.nanmedian(res.mean_latency[:500])
    # stable replicas: latency floor = 2 ms CPU + 12 ms IO, percentiles ordered
    assert (st["lat_min"][st["completed"] > 0] >= 0.014).all()
    ok = st["completed"] > 10
    assert (st["p50"][ok] <= st["p95"][ok]).all() and (st["p95"][ok] <= st["p99"][ok]).all()
    assert (st["p99"][ok] <= st["lat_max"][ok] * 1.02).all()
    # determinism: the same sweep again gives identical bits
    again = sw.run()
    np.testing.assert_array_equal(again.stats, st)
    sw.close()


def test_c3_statistics_match_reference_dashboard():
    """configs[2] scenario at the reference's own parameters: the README dashboard reads
    mean 0.024 / p50 0.023 / p95 0.034 / p99 0.040 s, 123.6 rps."""
    flat = flatten(load_scenario("c3_lb_two_servers.yml", 600))
    sw = SweepRunner(flat, 296, seed=SEED, throughput=True)
    res = sw.run()
    s = res.summary()
    assert abs(s["mean_latency"] - 0.024) < 0.0008
    assert abs(s["p50_mean"] - 0.023) < 0.0008
    assert abs(s["p95_mean"] - 0.034) < 0.0012
    assert abs(s["p99_mean"] - 0.040) < 0.0016
    rps = res.completed.mean() / 600
    assert abs(rps - 133.3 * 0.99 ** 4) < 1.5
    # round robin keeps the two LB edges within one request of each other (no outages)
    a, b = flat.edge_ids.index("lb-srv1"), flat.edge_ids.index("lb-srv2")
    assert (np.abs(res.edge_sent[:, a].astype(np.int64) - res.edge_sent[:, b]) <= 1).all()
    sw.close()


@pytest.mark.parametrize("seed", [101, 107, 111, 113, 118, 122])
def test_engine_equals_oracle_on_random_tie_prone_scenarios(eng, seed):
    """tests/fuzz.py: deterministic step durations that tie, CPU/RAM queueing, every distribution."""
    import fuzz
    payload = fuzz.scenario(seed)
    flat = flatten(payload)
    eng.upload(flat)
    eng.configure(trace_replicas=1, trace_clock_capacity=100000, request_capacity=200000)
    eng.run(SEED, seed, seed + 1)
    st = eng.stats()
    sent, dropped = eng.edge_counts()
    assert st[0]["flags"] == 0
    o = des_port.simulate(payload, seed=SEED, replica=seed)
    assert_matches_oracle(o, flat, stats=st[0], clocks=eng.trace_clocks(0), sent=sent[0], dropped=dropped[0],
                          series=eng.trace_series(0), throughput=eng.throughput()[0])


def test_run_sharded_world_of_one_equals_plain_run():
    from asyncflow_b200.distributed import run_sharded
    flat = flatten(load_scenario("c1_my_service.yml", 8))
    sw = SweepRunner(flat, 64, {("users_mean",): np.linspace(20, 200, 64)}, seed=SEED)
    res, glob = run_sharded(sw)
    assert glob.replicas == 64 and glob.completed == int(res.completed.sum())
    assert int(glob.histogram.sum()) == glob.completed
    lat_all = glob.percentile(95)
    assert res.stats["p95"].min() <= lat_all <= res.stats["p95"].max()
    sw.close()


def test_two_engines_on_one_device_do_not_disturb_each_other():
    """Launch parameters are in __constant__ memory (per device): a second engine has to wait."""
    fa, fb = flatten(load_scenario("c1_my_service.yml", 10)), flatten(load_scenario("c3_lb_two_servers.yml", 8))
    with Engine(0) as a, Engine(0) as b:
        a.upload(fa); b.upload(fb)
        a.configure(); b.configure()
        a.run(SEED, 0, 2000); b.run(SEED, 0, 2000)      # b is launched while a may still be running
        sa, sb = a.stats().copy(), b.stats().copy()
        a.run(SEED, 0, 2000); a.sync(); b.run(SEED, 0, 2000); b.sync()
        np.testing.assert_array_equal(a.stats(), sa)
        np.testing.assert_array_equal(b.stats(), sb)
    assert sa["completed"].sum() > 0 and sb["completed"].sum() > 0


def test_the_timed_workload_is_pinned_to_the_oracle(eng, mode):
    """bench.py's exact workload (configs[2]: C3, every edge normal(mean = RTT, sigma = jitter x RTT), the row a
    GLOBAL replica id gets from bench.sweep_rows, horizon 60 s): first, middle and last replica of the 100 000,
    every (start, finish), counter, per-second bucket and sampled series against the oracle on the payload the
    sweep row stands for."""
    import bench
    total = 100_000
    payload = bench.workload(total, 60)
    flat = flatten(payload)
    ids = np.arange(total, dtype=np.int64)
    rtt, sig = bench.sweep_rows(ids, total)
    cols = {}
    for e in flat.edge_ids:
        cols[("edge_mean", e)] = rtt
        cols[("edge_sigma", e)] = sig
    spec = SweepSpec(flat, total, cols)
    eng.upload(flat)
    for rid in (0, total // 2 - 1, total - 1):
        eng.configure(trace_replicas=1, trace_clock_capacity=20000, throughput=True, **MODES[mode])
        eng.upload_sweep(spec, rid, row_first=rid, row_count=1)
        eng.run(bench.SEED, rid, rid + 1)
        st = eng.stats()
        sent, dropped = eng.edge_counts()
        assert st[0]["flags"] == 0
        o = des_port.simulate(spec.payload_for(payload, rid), seed=bench.SEED, replica=rid)
        assert o["completed"] > 7000
        assert_matches_oracle(o, flat, stats=st[0], clocks=eng.trace_clocks(0), sent=sent[0], dropped=dropped[0],
                              series=eng.trace_series(0), throughput=eng.throughput()[0], hist=eng.histograms()[0])
    eng.upload_sweep(None)


def test_every_bench_config_row_matches_the_oracle(eng):
    """One sweep row of each bench.py --config (c2, c4, c5 at their BASELINE shapes, shortened horizons): the
    engine in its product mode against the oracle on SweepSpec.payload_for(row)."""
    import bench
    for cfg, horizon, rows in (("c2", 20, (0, 9_999)), ("c4", 125, (123_456,)), ("c5", 6, (3_999_999,))):
        w = bench.make_workload(cfg, horizon=horizon)
        flat = flatten(w.payload)
        spec = SweepSpec(flat, len(rows), w.columns(flat, np.asarray(rows, dtype=np.int64), w.replicas))
        eng.upload(flat)
        for i, rid in enumerate(rows):
            eng.configure(trace_replicas=1, trace_clock_capacity=400000, request_capacity=400000, throughput=True)
            eng.upload_sweep(spec, rid, row_first=i, row_count=1)
            eng.run(bench.SEED, rid, rid + 1)
            st = eng.stats()
            sent, dropped = eng.edge_counts()
            assert st[0]["flags"] == 0, (cfg, rid, int(st[0]["flags"]))
            o = des_port.simulate(spec.payload_for(w.payload, i), seed=bench.SEED, replica=rid)
            assert_matches_oracle(o, flat, stats=st[0], clocks=eng.trace_clocks(0), sent=sent[0], dropped=dropped[0],
                                  series=eng.trace_series(0), throughput=eng.throughput()[0])
    eng.upload_sweep(None)


def test_from_yaml_on_the_device(eng, tmp_path):
    """GpuSimulationRunner.from_yaml(env=, yaml_path=): reference runtime/simulation_runner.py:381-398."""
    import yaml
    payload = load_scenario("c3_lb_two_servers.yml", 10)
    path = tmp_path / "two_servers_lb.yml"
    path.write_text(yaml.safe_dump(payload))
    res = GpuSimulationRunner.from_yaml(env=None, yaml_path=path, seed=SEED, replica=2).run()
    o = des_port.simulate(payload, seed=SEED, replica=2)
    np.testing.assert_array_equal(res.clocks, np.array(o["clocks"]).reshape(-1, 2))
    assert res.get_latency_stats()["total_requests"] == o["completed"]


def test_flagged_replicas_are_rerun_inside_af_run(eng):
    """TWO_PASS mode (AUTO for large launches): the thread-per-replica pass has nominal-load pools; the saturated rows of a users sweep overflow
    them and are re-run one per warp inside the same af_run -- results complete, flags clear, rows exact."""
    base = load_scenario("c1_my_service.yml", 20)
    flat = flatten(base)
    users = [30.0, 900.0, 40.0, 1000.0, 850.0, 20.0, 100.0, 700.0]
    spec = SweepSpec(flat, len(users), {("users_mean",): users})
    eng.set_mode("two_pass")
    eng.upload(flat)
    eng.configure(trace_replicas=len(users), trace_clock_capacity=40000, request_capacity=40000, throughput=True)
    eng.upload_sweep(spec, 0)
    eng.run(SEED, 0, len(users))
    st = eng.stats()
    sent, dropped = eng.edge_counts()
    passes = eng.last_run_passes()
    assert passes["lane_replicas"] == len(users) and 3 <= passes["warp_replicas"] <= 4, passes
    assert (st["flags"] == 0).all()
    assert int(st["peak_requests"].max()) > 2048
    for i in range(len(users)):
        o = des_port.simulate(spec.payload_for(base, i), seed=SEED, replica=i)
        assert_matches_oracle(o, flat, stats=st[i], clocks=eng.trace_clocks(i), sent=sent[i], dropped=dropped[i],
                              series=eng.trace_series(i), throughput=eng.throughput()[i], hist=eng.histograms()[i])
    eng.upload_sweep(None)


def test_auto_mode_picks_the_kernel_by_launch_size_and_the_result_does_not_care(eng):
    """AF_MODE_AUTO: one replica per GPU thread from 3 x SMs x 32 replicas up, one per warp below (a thread advances one
    replica ~10x slower than a warp: it pays off in numbers) -- and every statistic of every replica is the same."""
    flat = flatten(load_scenario("c1_my_service.yml", 3))
    n = 20000

    def run(mode, count):
        eng.set_mode(mode)
        eng.upload(flat)
        eng.configure(throughput=False)
        eng.run(SEED, 0, count)
        return eng.stats().copy(), eng.last_run_passes()

    try:
        small, p_small = run("auto", 64)
        big, p_big = run("auto", n)
        warp, _ = run("warp", n)
    finally:
        eng.set_mode("auto")
    assert p_small["lane_pass"] == 0 and p_small["warp_pass"] == 1
    assert p_big["lane_pass"] == 1 and p_big["lane_replicas"] == n and 4 <= p_big["lane_warps_per_sm"] <= 12
    for f in ("n_events", "generated", "completed", "flags", "n_ticks", "lat_sum", "lat_sumsq", "lat_min", "lat_max", "p50", "p95", "p99"):
        np.testing.assert_array_equal(big[f], warp[f], err_msg=f)
        np.testing.assert_array_equal(small[f], warp[f][:64], err_msg=f)


def test_exact_percentiles_on_the_device():
    """SweepRunner.exact_percentiles: rows re-simulated with their clock lists kept, numpy.percentile on them == the
    oracle's latencies' percentiles, bit for bit; the other rows keep the histogram values (within 1 %)."""
    from asyncflow_b200 import SweepRunner
    base = load_scenario("c1_my_service.yml", 6)
    n = 300
    users = np.linspace(20.0, 200.0, n)
    sw = SweepRunner(base, n, {("users_mean",): users}, seed=SEED)
    try:
        res = sw.run()
        coarse = res.stats[["p50", "p95", "p99"]].copy()
        rows = [0, 1, 150, 299]
        sw.exact_percentiles(res, rows, chunk=2)
        for row in rows + [7]:
            o = des_port.simulate(sw.payload_for(row), seed=SEED, replica=row)
            lat = np.array([b - a for a, b in o["clocks"]])
            for q, key in ((50, "p50"), (95, "p95"), (99, "p99")):
                exact = float(np.percentile(lat, q))
                if row in rows:
                    assert float(res.stats[key][row]) == exact, (row, key)
                else:
                    assert float(res.stats[key][row]) == float(coarse[key][row])
                    assert abs(float(res.stats[key][row]) - exact) <= 0.01 * exact
    finally:
        sw.close()

