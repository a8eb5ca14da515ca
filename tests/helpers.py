"""Helpers shared by the CPU-tier and GPU-tier tests."""

from __future__ import annotations

import hashlib
import json
from pathlib import Path

import numpy as np
import yaml

ROOT = Path(__file__).resolve().parent.parent
SCEN = ROOT / "tests" / "scenarios"
GOLD = ROOT / "tests" / "golden"
SEED = 0xA5F10
SERVER_SERIES = ("ready_queue_len", "event_loop_io_sleep", "ram_in_use")

#: scenario -> horizon used in the oracle-vs-engine parity tests (None = as written)
PARITY_CASES = {
    "c1_my_service.yml": 20, "c3_lb_two_servers.yml": 30, "c4_lb8_events.yml": 250,
    "ev_spikes_outages.yml": None, "mixed_lc.yml": None, "overload_single.yml": None,
    "chain_two_servers.yml": None, "poisson_ties.yml": None, "tie_cpu_io.yml": None,
    "c5_multihop32.yml": 8,
}


def load_scenario(name: str, horizon: int | None = None) -> dict:
    d = yaml.safe_load((SCEN / name).read_text())
    if horizon is not None:
        d["sim_settings"]["total_simulation_time"] = horizon
    return d


def load_golden(name: str) -> dict:
    return json.loads((GOLD / (Path(name).stem + ".json")).read_text())


def sha(arr: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def unhex(pairs) -> np.ndarray:
    return np.array([[float.fromhex(a), float.fromhex(b)] for a, b in pairs], dtype=np.float64).reshape(-1, 2)


def check_against_golden(vec: dict, *, generated, completed, clocks, edge_sent, edge_dropped,
                         throughput=None, series=None, flat=None) -> None:
    """`clocks` is an [n,2] f64 array in completion order; `series` [n_series, n_ticks] u32."""
    assert generated == vec["generated"]
    assert completed == vec["completed"]
    assert dict(edge_sent) == vec["edge_sent"]
    assert dict(edge_dropped) == vec["edge_dropped"]
    clocks = np.ascontiguousarray(clocks, dtype="<f8").reshape(-1, 2)
    assert clocks.shape[0] == vec["completed"]
    np.testing.assert_array_equal(clocks[:32], unhex(vec["clocks_head"]))
    np.testing.assert_array_equal(clocks[-32:], unhex(vec["clocks_tail"]))
    assert sha(clocks) == vec["clocks_sha256"]
    if "clocks" in vec:
        np.testing.assert_array_equal(clocks, unhex(vec["clocks"]))
    if throughput is not None:
        assert [int(x) for x in throughput] == vec["throughput"]
    if series is not None:
        for si, sid in enumerate(flat.server_ids):
            for mi, m in enumerate(SERVER_SERIES):
                g = vec["server_series"][sid].get(m)
                if g is None:
                    continue
                row = np.ascontiguousarray(series[3 * si + mi], dtype="<u4")
                assert len(row) == g["n"] and int(row.sum()) == g["sum"] and sha(row) == g["sha256"], (sid, m)
        for ei, eid in enumerate(flat.edge_ids):
            g = vec["edge_series"][eid].get("edge_concurrent_connection")
            if g is None:
                continue
            row = np.ascontiguousarray(series[3 * flat.n_servers + ei], dtype="<u4")
            assert len(row) == g["n"] and int(row.sum()) == g["sum"] and sha(row) == g["sha256"], eid


def oracle_series_matrix(o: dict, flat) -> np.ndarray:
    """Oracle sampled series in the engine's row order, [n_series, n_ticks]."""
    rows = []
    for sid in flat.server_ids:
        for m in SERVER_SERIES:
            rows.append(o["server_series"][sid].get(m, []))
    for eid in flat.edge_ids:
        rows.append(o["edge_series"][eid].get("edge_concurrent_connection", []))
    n = max((len(r) for r in rows), default=0)
    out = np.zeros((len(rows), n), dtype=np.uint32)
    for i, r in enumerate(rows):
        out[i, : len(r)] = r
    return out


def assert_matches_oracle(o: dict, flat, *, stats, clocks, sent, dropped, series=None,
                          throughput=None, hist=None) -> None:
    """Bit-exact comparison of one engine replica with one oracle replica."""
    assert int(stats["generated"]) == o["generated"]
    assert int(stats["completed"]) == o["completed"]
    assert [int(x) for x in sent] == [o["edge_sent"][e] for e in flat.edge_ids]
    assert [int(x) for x in dropped] == [o["edge_dropped"][e] for e in flat.edge_ids]
    oc = np.array(o["clocks"], dtype=np.float64).reshape(-1, 2)
    if clocks is not None:
        np.testing.assert_array_equal(np.asarray(clocks).reshape(-1, 2), oc)
    lat = oc[:, 1] - oc[:, 0]
    # sequential sums in completion order: the engine accumulates in the same order
    s = 0.0
    s2 = 0.0
    for x in lat.tolist():
        s += x
        s2 += x * x
    assert float(stats["lat_sum"]) == s
    assert float(stats["lat_sumsq"]) == s2
    if len(lat):
        assert float(stats["lat_min"]) == float(lat.min())
        assert float(stats["lat_max"]) == float(lat.max())
    if series is not None:
        np.testing.assert_array_equal(np.asarray(series), oracle_series_matrix(o, flat))
    if throughput is not None:
        thr = np.zeros(flat.horizon_s, dtype=np.int64)
        for f in oc[:, 1]:
            thr[int(np.ceil(f)) - 1] += 1
        np.testing.assert_array_equal(np.asarray(throughput, dtype=np.int64), thr)
    if hist is not None:
        assert int(np.asarray(hist).sum()) == o["completed"]
        if len(lat):
            # percentiles read off the histogram (128 bins per octave: <= 0.78 % wide) are within 1 % of numpy's exact
            # ones -- half of the 2 % the north star allows
            for q, key in ((50, "p50"), (95, "p95"), (99, "p99")):
                exact = float(np.percentile(lat, q))
                assert abs(float(stats[key]) - exact) <= 0.01 * exact, (key, float(stats[key]), exact)
