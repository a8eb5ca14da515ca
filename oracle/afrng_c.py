"""ctypes face of ``oracle/afrng_c.c`` -- ORACLE / TEST INFRASTRUCTURE ONLY.

Same numbers as ``oracle/afrng.py`` (checked bit-for-bit by tests/test_afrng.py);
used by ``oracle/des_port.py`` so a draw costs what numpy costs the reference.
"""

from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libafrng_c.so"


def build(force: bool = False) -> Path:
    src = _HERE / "afrng_c.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    return _SO


def _load() -> C.CDLL:
    build()
    lib = C.CDLL(str(_SO))
    u64, u32, dbl, i64 = C.c_uint64, C.c_uint32, C.c_double, C.c_int64
    pu32, pdbl = C.POINTER(u32), C.POINTER(dbl)
    lib.afrng_philox.argtypes = [pu32, pu32, pu32]
    lib.afrng_philox.restype = None
    lib.afrng_log.argtypes = [dbl]
    lib.afrng_log.restype = dbl
    lib.afrng_exp.argtypes = [dbl]
    lib.afrng_exp.restype = dbl
    lib.afrng_gen_uniform.argtypes = [u64, u64, pu32]
    lib.afrng_gen_uniform.restype = dbl
    lib.afrng_gen_poisson.argtypes = [u64, u64, pu32, dbl]
    lib.afrng_gen_poisson.restype = i64
    lib.afrng_gen_normal.argtypes = [u64, u64, pu32, dbl, dbl]
    lib.afrng_gen_normal.restype = dbl
    lib.afrng_edge.argtypes = [u64, u64, u32, u32, C.c_int, dbl, dbl, pdbl, pdbl]
    lib.afrng_edge.restype = None
    lib.afrng_endpoint.argtypes = [u64, u64, u32, u32, u32]
    lib.afrng_endpoint.restype = u32
    return lib


LIB = _load()


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    LIB.afrng_philox(c, k, o)
    return tuple(o)


class CRng:
    """Same interface as ``des_port.PyRng``."""

    def __init__(self, seed: int, replica: int) -> None:
        self.seed, self.replica = seed, replica
        self._pos = C.c_uint32(0)
        self._ppos = C.byref(self._pos)
        self._u = C.c_double()
        self._l = C.c_double()
        self._pu = C.byref(self._u)
        self._pl = C.byref(self._l)

    def gen_uniform(self) -> float:
        return LIB.afrng_gen_uniform(self.seed, self.replica, self._ppos)

    def gen_poisson(self, lam: float) -> int:
        return LIB.afrng_gen_poisson(self.seed, self.replica, self._ppos, lam)

    def gen_normal(self, mean: float, sigma: float) -> float:
        return LIB.afrng_gen_normal(self.seed, self.replica, self._ppos, mean, sigma)

    def edge(self, rid: int, hop: int, dist: int, mean: float, sigma: float):
        LIB.afrng_edge(self.seed, self.replica, rid, hop, dist, mean, sigma, self._pu, self._pl)
        return self._u.value, self._l.value

    def endpoint(self, rid: int, hop: int, n: int) -> int:
        return LIB.afrng_endpoint(self.seed, self.replica, rid, hop, n)
