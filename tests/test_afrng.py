"""AF-RNG spec: Philox known answers, C restatement == Python spec bit for bit."""

from __future__ import annotations

import math
import random
import struct

import afrng
import afrng_c
import pytest


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    for ctr, key, out in afrng.PHILOX_KAT:
        assert afrng.philox4x32_10(ctr, key) == out
        assert afrng_c.philox(ctr, key) == out


def _ulps(a: float, b: float) -> int:
    ia = struct.unpack("<q", struct.pack("<d", a))[0]
    ib = struct.unpack("<q", struct.pack("<d", b))[0]
    return abs(ia - ib)


def test_log_exp_match_c_and_are_within_one_ulp_of_libm():
    rnd = random.Random(7)
    worst_log = worst_exp = 0
    for _ in range(20000):
        x = rnd.random() * 10 ** rnd.uniform(-15, 0) or 0.5
        assert afrng.af_log(x) == afrng_c.LIB.afrng_log(x)
        worst_log = max(worst_log, _ulps(afrng.af_log(x), math.log(x)))
        y = rnd.uniform(-300, 300)
        assert afrng.af_exp(y) == afrng_c.LIB.afrng_exp(y)
        worst_exp = max(worst_exp, _ulps(afrng.af_exp(y), math.exp(y)))
    assert worst_log <= 1 and worst_exp <= 1
    assert afrng.af_log(1.0) == 0.0


@pytest.mark.parametrize("dist,mean,sigma", [
    (afrng.D_EXPONENTIAL, 0.003, 0.0), (afrng.D_NORMAL, 0.02, 0.01), (afrng.D_NORMAL, 0.001, 0.01),
    (afrng.D_LOG_NORMAL, 0.001, 0.25), (afrng.D_UNIFORM, 0.5, 0.0), (afrng.D_POISSON, 0.4, 0.0),
    (afrng.D_POISSON, 3.5, 0.0),
])
def test_edge_draws_c_equals_python(dist, mean, sigma):
    seed, rep = 0xA5F10, 123456789012
    c = afrng_c.CRng(seed, rep)
    for rid in range(1, 400):
        for hop in (1, 3, 5, 7):
            d = afrng.RequestDraw(seed, rep, afrng.P_EDGE, rid, hop)
            want = (d.head53(), afrng.sample_rv(dist, mean, sigma, d))
            assert c.edge(rid, hop, dist, mean, sigma) == want
            assert c.endpoint(rid, hop, 7) == afrng.pick_endpoint(seed, rep, rid, hop, 7)


def test_generator_stream_c_equals_python():
    seed, rep = 99, 3
    g = afrng.GenStream(seed, rep)
    c = afrng_c.CRng(seed, rep)
    for i in range(300):
        if i % 50 == 0:
            assert c.gen_poisson(700.0) == afrng.poisson(700.0, g)
            assert c.gen_poisson(0.7) == afrng.poisson(0.7, g)
            assert c.gen_normal(50.0, 12.0) == 50.0 + 12.0 * afrng.std_normal(g)
        assert c.gen_uniform() == g.next53()


def test_variates_have_the_right_moments():
    g = afrng.GenStream(1, 2)
    n = 4000
    xs = [afrng.poisson(100.0, g) for _ in range(n)]
    m = sum(xs) / n
    v = sum((x - m) ** 2 for x in xs) / n
    assert abs(m - 100.0) < 1.0 and abs(v - 100.0) < 10.0
    zs = [afrng.std_normal(g) for _ in range(n)]
    mz = sum(zs) / n
    vz = sum((z - mz) ** 2 for z in zs) / n
    assert abs(mz) < 0.06 and abs(vz - 1.0) < 0.08
    es = [afrng.std_exponential(g) for _ in range(n)]
    assert abs(sum(es) / n - 1.0) < 0.06
