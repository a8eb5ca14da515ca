"""A stand-in for ``asyncflow_b200.engine.Engine`` backed by the CPU twin -- TEST INFRASTRUCTURE ONLY.

Same Python surface as the ctypes engine (configure / upload / upload_sweep / run / fetchers), so the
host-side orchestration in ``SweepRunner`` (launch order, shard ranges, traced replicas, result
un-permutation) can be exercised here without a GPU.  Never imported by the product package."""

from __future__ import annotations

import numpy as np
import twin

from asyncflow_b200 import _capi as K


class TwinEngine:
    def __init__(self, device: int = 0) -> None:
        self.device = device
        self.flat = None
        self.opt: dict = {}
        self.sweep = None
        self._n = 0
        self._r: dict = {}
        self.calls: list = []

    # -- set-up -----------------------------------------------------------------
    def configure(self, **kw) -> None:
        self.opt = kw
        self.calls.append(("configure", dict(kw)))

    def upload(self, flat) -> None:
        self.flat = flat

    def upload_sweep(self, spec, first_replica: int = 0, *, row_first: int = 0, row_count=None) -> None:
        assert first_replica == row_first, "the twin addresses sweep rows by replica id"
        self.sweep = spec
        self.calls.append(("upload_sweep", first_replica, row_first, row_count))

    def close(self) -> None:
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    # -- run --------------------------------------------------------------------
    def run(self, seed: int, begin: int, end: int) -> None:
        o = self.opt
        self._n = end - begin
        self._r = twin.run(self.flat, seed=seed, replica_begin=begin, n=self._n, sweep=self.sweep, sweep_first=begin,
                           trace=o.get("trace_replicas", 0), clock_cap=o.get("trace_clock_capacity", 0),
                           event_capacity=o.get("event_capacity", 0), request_capacity=o.get("request_capacity", 0))
        self.calls.append(("run", seed, begin, end))

    def sync(self) -> None:
        pass

    def last_run_ms(self):
        return (1.0, 1.0)

    def last_run_passes(self) -> dict:
        return {"lane_pass": 1, "warp_pass": 0, "lane_warps_per_sm": 0, "lane_bytes": twin.DEFAULT_LANE_BYTES,
                "lane_events_smem": 0, "lane_requests_smem": 0, "lane_replicas": self._n, "warp_replicas": 0}

    # -- fetchers ---------------------------------------------------------------
    def _into(self, out, arr):
        if out is None:
            return arr.copy()
        out[...] = arr
        return out

    def stats(self, out=None):
        return self._into(out, self._r["stats"])

    def edge_counts(self, sent=None, dropped=None):
        return self._into(sent, self._r["sent"]), self._into(dropped, self._r["dropped"])

    def sampled(self, sums=None, maxima=None):
        return self._into(sums, self._r["samp_sum"]), self._into(maxima, self._r["samp_max"])

    def throughput(self, out=None):
        return self._into(out, self._r["thr"])

    def histograms(self):
        return self._r["hist"].copy()

    def reduced_histogram(self):
        return self._r["hist"].astype(np.uint64).sum(axis=0)

    def trace_clocks(self, j: int):
        return self._r["trace_clocks"][j, : int(self._r["stats"][j]["completed"])].copy()

    def trace_series(self, j: int):
        return self._r["trace_series"][j][:, : int(self._r["stats"][j]["n_ticks"])].copy()


assert K.AF_HIST_BINS == 4096
