"""AF-RNG: the counter-based random-variate SPEC shared by oracle and device.

ORACLE / TEST INFRASTRUCTURE ONLY -- nothing under ``asyncflow_b200/`` imports
this file.  It is the *normative*, pure-Python statement of the random numbers
every replica consumes; ``oracle/afrng_c.c`` (host C, used for speed) and
``asyncflow_b200/csrc/af_rng.cuh`` (device) are independent restatements that
``tests/test_afrng.py`` / ``tests/test_gpu_rng.py`` check against it bit for
bit.

Why our own RNG: the reference draws from one *unseeded* ``numpy`` PCG64
shared by all actors (reference ``runtime/simulation_runner.py:77``); the only
supported seeding seam is assigning ``runner.rng`` / the actors' ``.rng`` with
a duck type (``tests/integration/single_server/test_int_single_server.py:36``,
``tests/unit/runtime/actors/test_edge.py:31-49``).  A sequential shared stream
cannot be consumed by 10^5 lock-step replicas, so the seam is used to inject
Philox4x32-10 (Salmon et al., SC'11 -- published algorithm and known-answer
vectors, see ``PHILOX_KAT``), keyed so that a draw depends only on
``(seed, replica, purpose, request id, hop)`` and never on event interleaving.

Counter / key layout (all u32):
    key = (seed_lo, seed_hi)
    ctr = (index, purpose<<24 | hop<<8 | block, replica_lo, replica_hi)

    purpose GEN   (0): index = block number j of the generator's sequential
                       uniform stream U_0,U_1,...; block j yields U_2j, U_2j+1.
                       (hop = block = 0)
    purpose EDGE  (1): index = request id, hop = len(history) when the request
                       enters the edge; block 0 words (0,1) -> dropout uniform,
                       words (2,3) -> first latency uniform / first polar
                       attempt; further uniforms come from blocks 1,2,...
    purpose SERVER(2): index = request id, hop = len(history) after the server
                       recorded its hop; word 0 -> endpoint pick.

All floating-point maths uses only IEEE-754 double +,-,*,/ and sqrt in a fixed
order (no fused multiply-add), so Python, gcc (-ffp-contract=off) and nvcc
(-fmad=false) produce identical bits.  ``af_log`` / ``af_exp`` follow the
published fdlibm (Sun/FreeBSD msun e_log.c, e_exp.c) algorithms, <1 ulp.
"""

from __future__ import annotations

import math
import struct

M32 = 0xFFFFFFFF
PHILOX_M0 = 0xD2511F53
PHILOX_M1 = 0xCD9E8D57
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85

P_GEN, P_EDGE, P_SERVER = 0, 1, 2

# distribution codes shared with the flattened scenario (include/asyncflow_b200.h)
D_POISSON, D_NORMAL, D_LOG_NORMAL, D_EXPONENTIAL, D_UNIFORM = 0, 1, 2, 3, 4
DIST_CODE = {
    "poisson": D_POISSON, "normal": D_NORMAL, "log_normal": D_LOG_NORMAL,
    "exponential": D_EXPONENTIAL, "uniform": D_UNIFORM,
}

#: Random123 known-answer vectors for philox4x32-10: (ctr, key) -> out
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0),
     (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((M32, M32, M32, M32), (M32, M32),
     (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def philox4x32_10(ctr, key):
    """Ten rounds of Philox-4x32.  ``ctr``: 4 u32, ``key``: 2 u32 -> 4 u32."""
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        c0, c1, c2, c3 = (
            ((p1 >> 32) ^ c1 ^ k0) & M32,
            p1 & M32,
            ((p0 >> 32) ^ c3 ^ k1) & M32,
            p0 & M32,
        )
        k0 = (k0 + PHILOX_W0) & M32
        k1 = (k1 + PHILOX_W1) & M32
    return c0, c1, c2, c3


def u53(hi_word: int, lo_word: int) -> float:
    """Two u32 words -> double in [0,1) with 53 random bits."""
    return ((hi_word >> 5) * 67108864 + (lo_word >> 6)) * (1.0 / 9007199254740992.0)


def s32(word: int) -> float:
    """One u32 word -> double in (-1,1), 32-bit resolution (polar attempts)."""
    return (word + 0.5) * (1.0 / 2147483648.0) - 1.0


# --------------------------------------------------------------------------- #
# deterministic elementary functions (fdlibm restatement, no FMA)             #
# --------------------------------------------------------------------------- #
_LN2_HI = 6.93147180369123816490e-01
_LN2_LO = 1.90821492927058770002e-10
_LG1 = 6.666666666666735130e-01
_LG2 = 3.999999999940941908e-01
_LG3 = 2.857142874366239149e-01
_LG4 = 2.222219843214978396e-01
_LG5 = 1.818357216161805012e-01
_LG6 = 1.531383769920937332e-01
_LG7 = 1.479819860511658591e-01


def _bits(x: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def _from_bits(b: int) -> float:
    return struct.unpack("<d", struct.pack("<Q", b & 0xFFFFFFFFFFFFFFFF))[0]


def af_log(x: float) -> float:
    """Natural log for finite normal ``x > 0`` (fdlibm e_log.c scheme)."""
    b = _bits(x)
    hx = b >> 32
    k = (hx >> 20) - 1023
    hx &= 0x000FFFFF
    i = (hx + 0x95F64) & 0x100000
    # normalise x or x/2 into [sqrt(2)/2, sqrt(2)]
    b = ((hx | (i ^ 0x3FF00000)) << 32) | (b & M32)
    k += i >> 20
    f = _from_bits(b) - 1.0
    dk = float(k)
    s = f / (2.0 + f)
    z = s * s
    w = z * z
    t1 = w * (_LG2 + w * (_LG4 + w * _LG6))
    t2 = z * (_LG1 + w * (_LG3 + w * (_LG5 + w * _LG7)))
    r = t2 + t1
    hfsq = 0.5 * f * f
    return dk * _LN2_HI - ((hfsq - (s * (hfsq + r) + dk * _LN2_LO)) - f)


_INVLN2 = 1.44269504088896338700e+00
_P1 = 1.66666666666666019037e-01
_P2 = -2.77777777770155933842e-03
_P3 = 6.61375632143793436117e-05
_P4 = -1.65339022054652515390e-06
_P5 = 4.13813679705723846039e-08


def af_exp(x: float) -> float:
    """exp(x) for |x| < 700 (fdlibm e_exp.c scheme)."""
    if x >= 0.0:
        k = int(_INVLN2 * x + 0.5)
    else:
        k = int(_INVLN2 * x - 0.5)  # truncation toward zero, like a C cast
    dk = float(k)
    hi = x - dk * _LN2_HI
    lo = dk * _LN2_LO
    r = hi - lo
    t = r * r
    c = r - t * (_P1 + t * (_P2 + t * (_P3 + t * (_P4 + t * _P5))))
    y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi)
    # scale by 2^k through the exponent field (y in [0.5, 2), |k| < 1020)
    return _from_bits(_bits(y) + (k << 52))


# --------------------------------------------------------------------------- #
# uniform sources                                                             #
# --------------------------------------------------------------------------- #
class UniformSource:
    """A well-defined sequence of uniforms for ONE variate draw.

    ``next53()`` -> next 53-bit uniform; ``next_pair32()`` -> next (v1, v2)
    polar attempt.  Sub-classes say where the words come from.
    """

    def next53(self) -> float:  # pragma: no cover - interface
        raise NotImplementedError

    def next_pair32(self):  # pragma: no cover - interface
        raise NotImplementedError


class GenStream(UniformSource):
    """Purpose GEN: strictly sequential U_0, U_1, ... for one replica."""

    def __init__(self, seed: int, replica: int) -> None:
        self.key = (seed & M32, (seed >> 32) & M32)
        self.rep = (replica & M32, (replica >> 32) & M32)
        self.i = 0  # index of the next uniform

    def next53(self) -> float:
        i = self.i
        self.i = i + 1
        w = philox4x32_10((i >> 1, P_GEN << 24, self.rep[0], self.rep[1]), self.key)
        return u53(w[0], w[1]) if (i & 1) == 0 else u53(w[2], w[3])

    def next_pair32(self):
        # the sequential stream has no cheap 32-bit lane: spend two uniforms
        a = self.next53()
        b = self.next53()
        return 2.0 * a - 1.0, 2.0 * b - 1.0


class RequestDraw(UniformSource):
    """Purposes EDGE / SERVER: the words belonging to (request, hop)."""

    def __init__(self, seed: int, replica: int, purpose: int, rid: int, hop: int) -> None:
        self.key = (seed & M32, (seed >> 32) & M32)
        self.c2 = replica & M32
        self.c3 = (replica >> 32) & M32
        self.rid = rid & M32
        self.tag = (purpose << 24) | ((hop & 0xFFFF) << 8)
        self._blk = {}
        self.pos = 2  # next free 64-bit half-block: block 0 words (2,3)

    def block(self, b: int):
        w = self._blk.get(b)
        if w is None:
            w = philox4x32_10((self.rid, self.tag | (b & 0xFF), self.c2, self.c3), self.key)
            self._blk[b] = w
        return w

    def head53(self) -> float:
        """Block 0 words (0,1): the dropout uniform."""
        w = self.block(0)
        return u53(w[0], w[1])

    def _next_words(self):
        p = self.pos
        self.pos = p + 2
        w = self.block(p >> 2)
        return w[p & 3], w[(p & 3) + 1]

    def next53(self) -> float:
        a, b = self._next_words()
        return u53(a, b)

    def next_pair32(self):
        a, b = self._next_words()
        return s32(a), s32(b)


# --------------------------------------------------------------------------- #
# variates                                                                    #
# --------------------------------------------------------------------------- #
def std_exponential(src: UniformSource) -> float:
    return -af_log(1.0 - src.next53())


def std_normal(src: UniformSource) -> float:
    """Marsaglia polar method (only log, sqrt, /): deterministic rejection."""
    while True:
        v1, v2 = src.next_pair32()
        s = v1 * v1 + v2 * v2
        if 0.0 < s < 1.0:
            return v1 * math.sqrt(-2.0 * af_log(s) / s)


POISSON_CHUNK = 256.0


def poisson(lam: float, src: UniformSource) -> int:
    """Exact Poisson(lam) by multiplication of uniforms, in chunks of <=256."""
    n = 0
    rem = lam
    while rem > 0.0:
        c = rem if rem < POISSON_CHUNK else POISSON_CHUNK
        rem = rem - c
        limit = af_exp(-c)
        p = 1.0
        while True:
            p = p * (1.0 - src.next53())
            if p <= limit:
                break
            n += 1
    return n


def sample_rv(dist: int, mean: float, sigma: float, src: UniformSource) -> float:
    """``general_sampler`` (reference ``samplers/common_helpers.py:49-89``).

    ``sigma`` is the schema's ``variance`` field, which numpy receives as the
    *standard deviation* (reference ``common_helpers.py:31,40``).
    """
    if dist == D_UNIFORM:
        return src.next53()  # U(0,1): the mean is ignored (common_helpers.py:62-65)
    if dist == D_POISSON:
        return float(poisson(mean, src))
    if dist == D_EXPONENTIAL:
        return mean * std_exponential(src)
    if dist == D_NORMAL:
        v = mean + sigma * std_normal(src)
        return v if v > 0.0 else 0.0  # max(0.0, value), common_helpers.py:32
    if dist == D_LOG_NORMAL:
        return af_exp(mean + sigma * std_normal(src))
    raise ValueError(f"unsupported distribution code {dist}")


def pick_endpoint(seed: int, replica: int, rid: int, hop: int, n: int) -> int:
    """``rng.integers(0, n)`` at reference ``runtime/actors/server.py:101``."""
    d = RequestDraw(seed, replica, P_SERVER, rid, hop)
    return (d.block(0)[0] * n) >> 32
