"""Randomised parity: the engine's state machine (CPU twin) vs the oracle, bit for bit, on
scenarios built to tie (deterministic step durations), queue (CPU and RAM) and branch."""

from __future__ import annotations

import des_port
import fuzz
import pytest
import twin
from helpers import SEED, assert_matches_oracle

from asyncflow_b200.flatten import flatten


@pytest.mark.parametrize("seed", range(100, 170))
def test_twin_equals_oracle_on_random_scenarios(seed):
    payload = fuzz.scenario(seed)
    flat = flatten(payload)
    o = des_port.simulate(payload, seed=SEED, replica=seed)
    r = twin.run(flat, seed=SEED, replica_begin=seed, n=1, trace=1, clock_cap=100000, request_capacity=200000)
    st = r["stats"][0]
    n, nt = int(st["completed"]), int(st["n_ticks"])
    assert st["flags"] == 0
    assert_matches_oracle(o, flat, stats=st, clocks=r["trace_clocks"][0, :n], sent=r["sent"][0],
                          dropped=r["dropped"][0], series=r["trace_series"][0][:, :nt], throughput=r["thr"][0])


@pytest.mark.parametrize("seed", range(300, 330))
def test_twin_equals_oracle_on_random_sweeps(seed):
    """Every sweep field of the C ABI (AF_FIELD_*): replica i of a random sweep == the oracle run on
    the scenario that SweepSpec.payload_for(i) spells out."""
    from asyncflow_b200.flatten import SweepSpec

    payload = fuzz.scenario(seed)
    flat = flatten(payload)
    n = 3
    spec = SweepSpec(flat, n, fuzz.sweep_columns(seed, payload, n))
    r = twin.run(flat, seed=SEED, replica_begin=0, n=n, sweep=spec, trace=n, clock_cap=100000,
                 request_capacity=200000)
    for i in range(n):
        p = spec.payload_for(payload, i)
        o = des_port.simulate(p, seed=SEED, replica=i)
        st = r["stats"][i]
        k, nt = int(st["completed"]), int(st["n_ticks"])
        assert st["flags"] == 0
        assert_matches_oracle(o, flatten(p), stats=st, clocks=r["trace_clocks"][i, :k], sent=r["sent"][i],
                              dropped=r["dropped"][i], series=r["trace_series"][i][:, :nt], throughput=r["thr"][i])


@pytest.mark.parametrize("seed", range(0, 16))
def test_twin_equals_oracle_on_big_topologies(seed):
    """C5-shaped: LB over 5-12 front ends chaining into shared back ends, overloaded -- hundreds to
    thousands of pending events and queued requests (the HBM tiers of both pools)."""
    payload = fuzz.big_scenario(seed)
    flat = flatten(payload)
    o = des_port.simulate(payload, seed=SEED, replica=seed)
    r = twin.run(flat, seed=SEED, replica_begin=seed, n=1, trace=1, clock_cap=200000, request_capacity=400000,
                 event_capacity=8192)
    st = r["stats"][0]
    n, nt = int(st["completed"]), int(st["n_ticks"])
    assert st["flags"] == 0
    assert_matches_oracle(o, flat, stats=st, clocks=r["trace_clocks"][0, :n], sent=r["sent"][0],
                          dropped=r["dropped"][0], series=r["trace_series"][0][:, :nt], throughput=r["thr"][0])
