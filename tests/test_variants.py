"""Build variants of the engine core (AF_PREDRAW / AF_PREGEN / AF_SORTED_POOL / AF_PIN_ACTIVE, af_core.cuh) must be BIT-IDENTICAL to the
product build: they only move where random numbers are computed (lane-parallel, memoised), never which
numbers.  Checked here on the CPU twin, byte for byte against the default twin (which the rest of the
suite pins to the oracle); tools/check_variant_gpu.py repeats it on the device."""

from __future__ import annotations

import ctypes as C

import fuzz
import numpy as np
import pytest
import twin
from helpers import PARITY_CASES, SEED, load_scenario

from asyncflow_b200.flatten import SweepSpec, flatten

VARIANTS = ["predraw", "pregen", "memo", "sorted", "pin", "all", "all4", "narrow", "tiny"]


def same(a: dict, b: dict) -> None:
    for k in a:
        if k == "stats":
            assert a[k].tobytes() == b[k].tobytes(), k
        else:
            np.testing.assert_array_equal(a[k], b[k], err_msg=k)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("name", sorted(PARITY_CASES))
def test_variant_equals_product_build_on_parity_scenarios(variant, name):
    flat = flatten(load_scenario(name, PARITY_CASES[name]))
    kw = dict(seed=SEED, replica_begin=3, n=2, trace=2, clock_cap=100000, request_capacity=200000)
    same(twin.run(flat, **kw), twin.run(flat, variant=variant, **kw))


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("seed", range(400, 430))
def test_variant_equals_product_build_on_random_scenarios_and_sweeps(variant, seed):
    payload = fuzz.scenario(seed)
    flat = flatten(payload)
    spec = SweepSpec(flat, 2, fuzz.sweep_columns(seed, payload, 2))
    kw = dict(seed=SEED, replica_begin=0, n=2, sweep=spec, trace=2, clock_cap=100000, request_capacity=200000)
    same(twin.run(flat, **kw), twin.run(flat, variant=variant, **kw))


@pytest.mark.parametrize("variant", ["memo", "sorted", "pin", "all4"])
@pytest.mark.parametrize("seed", range(16, 28))
def test_variant_equals_product_build_on_big_topologies(variant, seed):
    flat = flatten(fuzz.big_scenario(seed))
    kw = dict(seed=SEED, replica_begin=seed, n=1, trace=1, clock_cap=200000, request_capacity=400000, event_capacity=8192)
    same(twin.run(flat, **kw), twin.run(flat, variant=variant, **kw))


def test_the_edge_memo_is_live_and_mostly_hits():
    L = twin.lib("predraw")
    out = (C.c_uint64 * 2)()
    for name, floor in (("c1_my_service.yml", 0.99), ("c3_lb_two_servers.yml", 0.99), ("c4_lb8_events.yml", 0.5)):
        flat = flatten(load_scenario(name, 10))
        L.af_twin_pre_lookups(out)
        twin.run(flat, seed=SEED, n=1, variant="predraw")
        L.af_twin_pre_lookups(out)
        miss, hit = int(out[0]), int(out[1])
        assert hit + miss > 1000 and hit / (hit + miss) >= floor, (name, miss, hit)


def test_the_sorted_pool_exercises_both_modes_and_both_switches():
    L = twin.lib("sorted")
    out = (C.c_uint64 * 4)()
    seen = np.zeros(4, dtype=np.int64)
    for name, horizon in (("c3_lb_two_servers.yml", 10), ("c4_lb8_events.yml", 120), ("poisson_ties.yml", None)):
        flat = flatten(load_scenario(name, horizon))
        L.af_twin_pool_counts(out)
        twin.run(flat, seed=SEED, n=1, variant="sorted", request_capacity=200000)
        L.af_twin_pool_counts(out)
        if name.startswith("c3"):
            assert out[1] == 0 and out[0] > 5000          # nominal load never leaves the sorted ring
        seen += np.array(list(out), dtype=np.int64)
    assert (seen > 0).all(), seen.tolist()                # ring pushes, unsorted pushes, A->B, B->A


def test_pinning_keeps_served_requests_in_the_fast_tier():
    """Saturated single server (thousands queued): share of request-record accesses that go to the HBM tier."""
    out = (C.c_uint64 * 2)()
    share = {}
    flat = flatten(load_scenario("overload_single.yml"))
    for variant in ("count", "pin_count"):
        L = twin.lib(variant)
        L.af_twin_tier_counts(out)
        r = twin.run(flat, seed=SEED, n=1, variant=variant, request_capacity=400000)
        L.af_twin_tier_counts(out)
        assert r["stats"][0]["peak_requests"] > 1000
        share[variant] = out[1] / (out[0] + out[1])
    assert share["count"] > 0.85 and share["pin_count"] < 0.25, share
