"""Thin object wrapper over the C ABI (``include/asyncflow_b200.h``).

Owns one ``af_engine`` (one CUDA device).  All arrays that cross the boundary
are caller-visible numpy arrays; device memory belongs to the C side.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as K
from .flatten import FlatScenario, SweepSpec


class EngineError(RuntimeError):
    pass


class Engine:
    def __init__(self, device: int = 0) -> None:
        self._lib = K.load()
        h = C.c_void_p()
        rc = self._lib.af_engine_create(int(device), C.byref(h))
        if rc != 0:
            msg = self._lib.af_last_error(None).decode()
            raise K.EngineUnavailable(f"af_engine_create({device}) failed ({rc}): {msg}")
        self._h = h
        self.device = int(device)
        self.flat: FlatScenario | None = None
        self._n = 0
        self._opt = K.AfOptions(0, 0, 0, 0, 1, 1, 0, 0)
        self._keep: list = []

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise EngineError(f"{what} failed ({rc}): {self._lib.af_last_error(self._h).decode()}")

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.af_engine_destroy(self._h)
            self._h = None

    def __del__(self) -> None:  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def __enter__(self) -> "Engine":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    # ------------------------------------------------------------------ inputs
    def configure(self, *, event_capacity: int = 0, request_capacity: int = 0,
                  warps_per_block: int = 0, blocks_per_sm: int = 0, histogram: bool = True,
                  throughput: bool = True, trace_replicas: int = 0,
                  trace_clock_capacity: int = 0) -> None:
        self._opt = K.AfOptions(event_capacity, request_capacity, warps_per_block, blocks_per_sm,
                                int(histogram), int(throughput), trace_replicas, trace_clock_capacity)
        self._check(self._lib.af_engine_configure(self._h, C.byref(self._opt)), "af_engine_configure")

    def set_mode(self, mode: str | int) -> None:
        """Pass structure of :meth:`run`: ``"two_pass"`` (thread-per-replica kernel, flagged replicas re-run one per
        warp), ``"auto"`` (two_pass when the launch is large enough for it to pay off, else one replica per warp),
        ``"warp"`` or ``"lane"`` (one kernel only) -- ``af_engine_set_mode``."""
        m = K.MODES[mode] if isinstance(mode, str) else int(mode)
        self._check(self._lib.af_engine_set_mode(self._h, m), "af_engine_set_mode")

    def upload(self, flat: FlatScenario) -> None:
        self._check(self._lib.af_scenario_upload(self._h, C.byref(flat.pod)), "af_scenario_upload")
        self.flat = flat

    def upload_sweep(self, spec: SweepSpec | None, first_replica: int = 0, *, row_first: int = 0,
                     row_count: int | None = None) -> None:
        """Rows ``[row_first, row_first+row_count)`` of ``spec`` describe replicas starting at ``first_replica``."""
        if spec is None:
            self._check(self._lib.af_sweep_upload(self._h, None, 0), "af_sweep_upload")
            return
        sw, keep = spec.pod(row_first, row_count)
        self._check(self._lib.af_sweep_upload(self._h, C.byref(sw), int(first_replica)), "af_sweep_upload")
        del keep

    # ------------------------------------------------------------------ run
    def run(self, seed: int, replica_begin: int, replica_end: int) -> None:
        self._check(self._lib.af_run(self._h, int(seed), int(replica_begin), int(replica_end)), "af_run")
        self._n = int(replica_end) - int(replica_begin)

    def sync(self) -> None:
        self._check(self._lib.af_sync(self._h), "af_sync")

    def last_run_ms(self) -> tuple[float, float]:
        a, b = C.c_float(), C.c_float()
        self._check(self._lib.af_last_run_ms(self._h, C.byref(a), C.byref(b)), "af_last_run_ms")
        return a.value, b.value

    def selftest_rng(self, kind: int, n: int, *, seed: int, replica: int, dist: int = 0, mean: float = 0.0,
                     sigma: float = 0.0, hop: int = 1) -> tuple[np.ndarray, np.ndarray]:
        """AF-RNG known-answer hook (``af_selftest_rng``): ``n`` device draws of ``kind`` as two f64 arrays."""
        a, b = np.empty(n, dtype=np.float64), np.empty(n, dtype=np.float64)
        self._check(self._lib.af_selftest_rng(self._h, int(seed), int(replica), int(kind), int(dist), float(mean),
                                              float(sigma), int(hop), int(n), a.ctypes.data, b.ctypes.data), "af_selftest_rng")
        return a, b

    def last_run_passes(self) -> dict:
        """What the last run did: which kernels, the lane pass's occupancy and pool sizes, replicas per pass."""
        p = K.AfRunPasses()
        self._check(self._lib.af_last_run_passes(self._h, C.byref(p)), "af_last_run_passes")
        return {k: int(getattr(p, k)) for k, _ in K.AfRunPasses._fields_}

    @property
    def launch_count(self) -> int:
        return int(self._lib.af_launch_count(self._h))

    # ------------------------------------------------------------------ outputs
    def stats(self, out: np.ndarray | None = None) -> np.ndarray:
        out = np.empty(self._n, dtype=K.STATS_DTYPE) if out is None else out
        self._check(self._lib.af_fetch_stats(self._h, out.ctypes.data, self._n), "af_fetch_stats")
        return out

    def edge_counts(self, sent: np.ndarray | None = None,
                    dropped: np.ndarray | None = None) -> tuple[np.ndarray, np.ndarray]:
        ne = self.flat.n_edges
        sent = np.empty((self._n, ne), dtype=np.uint32) if sent is None else sent
        dropped = np.empty((self._n, ne), dtype=np.uint32) if dropped is None else dropped
        self._check(self._lib.af_fetch_edge_counts(self._h, sent.ctypes.data, dropped.ctypes.data, self._n),
                    "af_fetch_edge_counts")
        return sent, dropped

    def histograms(self) -> np.ndarray:
        out = np.empty((self._n, K.AF_HIST_BINS), dtype=np.uint32)
        self._check(self._lib.af_fetch_histograms(self._h, out.ctypes.data, self._n), "af_fetch_histograms")
        return out

    def throughput(self, out: np.ndarray | None = None) -> np.ndarray:
        out = np.empty((self._n, self.flat.horizon_s), dtype=np.uint32) if out is None else out
        self._check(self._lib.af_fetch_throughput(self._h, out.ctypes.data, self._n), "af_fetch_throughput")
        return out

    def sampled(self, sums: np.ndarray | None = None,
                maxima: np.ndarray | None = None) -> tuple[np.ndarray, np.ndarray]:
        ns = self.flat.n_series
        sums = np.empty((self._n, ns), dtype=np.uint64) if sums is None else sums
        maxima = np.empty((self._n, ns), dtype=np.uint32) if maxima is None else maxima
        self._check(self._lib.af_fetch_sampled(self._h, sums.ctypes.data, maxima.ctypes.data, self._n),
                    "af_fetch_sampled")
        return sums, maxima

    def reduced_histogram(self) -> np.ndarray:
        """Sum of all replicas' latency histograms, reduced on the device (``[AF_HIST_BINS]`` u64)."""
        out = np.empty(K.AF_HIST_BINS, dtype=np.uint64)
        self._check(self._lib.af_reduce_histograms(self._h, out.ctypes.data), "af_reduce_histograms")
        return out

    def trace_clocks(self, local_replica: int) -> np.ndarray:
        cap = max(1, self._opt.trace_clock_capacity)
        out = np.empty((cap, 2), dtype=np.float64)
        n = C.c_uint64()
        self._check(self._lib.af_fetch_trace_clocks(self._h, int(local_replica), out.ctypes.data, cap, C.byref(n)),
                    "af_fetch_trace_clocks")
        return out[: n.value].copy()

    def trace_series(self, local_replica: int) -> np.ndarray:
        """``[n_series, n_ticks]`` sampled values (see ``af_fetch_sampled`` for the series order)."""
        flat = self.flat
        cap = int(flat.horizon_s / flat.sample_period) + 2
        out = np.empty((flat.n_series, cap), dtype=np.uint32)
        n = C.c_uint64()
        self._check(self._lib.af_fetch_trace_series(self._h, int(local_replica), out.ctypes.data, cap, C.byref(n)),
                    "af_fetch_trace_series")
        return out[:, : n.value].copy()
