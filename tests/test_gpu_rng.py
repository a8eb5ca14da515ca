"""GPU tier: AF-RNG on the device == oracle/afrng_c, bit for bit, OUTSIDE the state machine.

SURVEY.md section 7 step 2's gate: 10^6 draws per distribution, host vs device bitwise.  The device side is
the C ABI's known-answer hook ``af_selftest_rng`` (asyncflow_b200/csrc/af_rng.cuh called directly from a
kernel); the host side is the oracle's C restatement of oracle/afrng.py (itself pinned to the Random123
Philox known answers by tests/test_afrng.py).
"""

from __future__ import annotations

import ctypes as C

import afrng
import afrng_c
import numpy as np
import pytest

from asyncflow_b200 import _capi as K
from asyncflow_b200 import Engine

pytestmark = pytest.mark.gpu

SEED, REPLICA = 0xA5F10, 123456789012
N = 1_000_000


@pytest.fixture(scope="module")
def eng():
    with Engine(0) as e:
        yield e


def _host_edges(n, hop, dist, mean, sigma):
    u, lat = np.empty(n), np.empty(n)
    cu, cl = C.c_double(), C.c_double()
    pu, pl = C.byref(cu), C.byref(cl)
    f = afrng_c.LIB.afrng_edge
    for i in range(n):
        f(SEED, REPLICA, i + 1, hop, dist, mean, sigma, pu, pl)
        u[i] = cu.value
        lat[i] = cl.value
    return u, lat


@pytest.mark.parametrize("dist,mean,sigma,hop", [
    (afrng.D_EXPONENTIAL, 0.003, 0.0, 1), (afrng.D_NORMAL, 0.02, 0.006, 3), (afrng.D_NORMAL, 0.001, 0.01, 5),
    (afrng.D_LOG_NORMAL, -6.0, 0.25, 5), (afrng.D_UNIFORM, 0.5, 0.0, 7), (afrng.D_POISSON, 3.5, 0.0, 3),
])
def test_edge_variates_device_equals_host(eng, dist, mean, sigma, hop):
    """The dropout uniform and the latency variate of 10^6 requests, every distribution of
    samplers/common_helpers.py:49-89."""
    u, lat = eng.selftest_rng(K.SELFTEST_EDGE, N, seed=SEED, replica=REPLICA, dist=dist, mean=mean, sigma=sigma, hop=hop)
    hu, hl = _host_edges(N, hop, dist, mean, sigma)
    assert np.array_equal(u.view(np.uint64), hu.view(np.uint64))
    assert np.array_equal(lat.view(np.uint64), hl.view(np.uint64))


def test_generator_stream_device_equals_host(eng):
    """10^6 consecutive uniforms of the generator stream and their -ln(1 - u) (poisson_poisson.py:72-74)."""
    u, nl = eng.selftest_rng(K.SELFTEST_GEN_UNIFORM, N, seed=SEED, replica=REPLICA)
    pos = C.c_uint32(0)
    f, lg = afrng_c.LIB.afrng_gen_uniform, afrng_c.LIB.afrng_log
    hu, hl = np.empty(N), np.empty(N)
    for i in range(N):
        x = f(SEED, REPLICA, C.byref(pos))
        hu[i] = x
        hl[i] = -lg(1.0 - max(x, 1e-15))
    assert pos.value == N
    assert np.array_equal(u.view(np.uint64), hu.view(np.uint64))
    assert np.array_equal(nl.view(np.uint64), hl.view(np.uint64))


@pytest.mark.parametrize("dist,mean,sigma,n", [(afrng.D_POISSON, 100.0, 0.0, 20000), (afrng.D_POISSON, 400.0, 0.0, 5000),
                                                (afrng.D_NORMAL, 80.0, 25.0, 200000)])
def test_user_draws_device_equals_host(eng, dist, mean, sigma, n):
    """First window draw of n replicas: Poisson users (poisson_poisson.py:58) / truncated normal
    (gaussian_poisson.py:70), value and stream position after the draw."""
    v, p = eng.selftest_rng(K.SELFTEST_GEN_USERS, n, seed=SEED, replica=REPLICA, dist=dist, mean=mean, sigma=sigma)
    hv, hp = np.empty(n), np.empty(n)
    for i in range(n):
        pos = C.c_uint32(0)
        if dist == afrng.D_POISSON:
            hv[i] = float(afrng_c.LIB.afrng_gen_poisson(SEED, REPLICA + i, C.byref(pos), mean))
        else:
            hv[i] = max(0.0, afrng_c.LIB.afrng_gen_normal(SEED, REPLICA + i, C.byref(pos), mean, sigma))
        hp[i] = pos.value
    assert np.array_equal(v.view(np.uint64), hv.view(np.uint64))
    assert np.array_equal(p, hp)


def test_endpoint_pick_device_equals_host(eng):
    """rng.integers(0, n) of runtime/actors/server.py:101 for 10^6 requests."""
    a, _ = eng.selftest_rng(K.SELFTEST_ENDPOINT, N, seed=SEED, replica=REPLICA, dist=7, hop=5)
    f = afrng_c.LIB.afrng_endpoint
    h = np.fromiter((f(SEED, REPLICA, i + 1, 5, 7) for i in range(N)), dtype=np.float64, count=N)
    assert np.array_equal(a, h)
