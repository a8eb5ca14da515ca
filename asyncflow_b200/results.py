"""Results: the OUT side of the drop-in boundary.

The reference hands its run to ``ResultsAnalyzer(client=, servers=, edges=,
settings=)`` (reference ``metrics/analyzer.py:51-58``), which reads only
duck-typed attributes: ``client.rqs_clock`` (objects with ``.start`` /
``.finish``), ``server.server_config.id`` + ``server.enabled_metrics``,
``edge.edge_config.id`` + ``edge.enabled_metrics`` and
``settings.total_simulation_time`` / ``sample_period_s`` (``analyzer.py:77-140``;
proved by ``tests/unit/metrics/test_analyzer.py:34-126``).

* :class:`ReplicaResults` carries the full trace of ONE materialised replica and
  answers the analyzer's getters itself (same keys, same arithmetic), and
  :meth:`ReplicaResults.to_reference_analyzer` builds the reference's *real*
  ``ResultsAnalyzer`` from holders when the ``asyncflow`` package is importable,
  so its matplotlib plots work unchanged.
* :class:`SweepResults` is the Monte-Carlo summary over many replicas
  (per-replica counters, exact mean/std/min/max, histogram percentiles).
"""

from __future__ import annotations

from collections import defaultdict
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Any

import numpy as np

from . import _capi as K
from .flatten import FlatScenario

SERVER_SERIES = ("ready_queue_len", "event_loop_io_sleep", "ram_in_use")
EDGE_SERIES = "edge_concurrent_connection"
LATENCY_ORDER = ("total_requests", "mean", "median", "std_dev", "p95", "p99", "min", "max")


class _Name(str):
    """A metric name that also answers ``.value`` like the reference's StrEnum."""

    @property
    def value(self) -> str:
        return str(self)


@dataclass
class ReplicaResults:
    flat: FlatScenario
    clocks: np.ndarray                 # [n, 2] (start, finish), completion order
    series: np.ndarray                 # [n_series, n_ticks]
    generated: int
    edge_sent: dict[str, int]
    edge_dropped: dict[str, int]
    n_events: int
    flags: int

    # -- analyzer-compatible getters (metrics/analyzer.py:147-262) -------------
    def list_server_ids(self) -> list[str]:
        return list(self.flat.server_ids)

    @property
    def latencies(self) -> np.ndarray:
        return self.clocks[:, 1] - self.clocks[:, 0]

    def get_latency_stats(self) -> dict[str, float]:
        arr = self.latencies
        if arr.size == 0:
            return {}
        return {
            "total_requests": float(arr.size), "mean": float(np.mean(arr)),
            "median": float(np.median(arr)), "std_dev": float(np.std(arr)),
            "p95": float(np.percentile(arr, 95)), "p99": float(np.percentile(arr, 99)),
            "min": float(np.min(arr)), "max": float(np.max(arr)),
        }

    def format_latency_stats(self) -> str:
        st = self.get_latency_stats()
        if not st:
            return "Latency stats: (empty)"
        lines = ["════════ LATENCY STATS ════════"]
        lines += [f"{k.upper():<20} = {st[k]:.6f}" for k in LATENCY_ORDER]
        return "\n".join(lines)

    def get_throughput_series(self, window_s: float | None = None) -> tuple[list[float], list[float]]:
        w = 1.0 if window_s is None else float(window_s)
        finish = np.sort(self.clocks[:, 1])
        end_time = self.flat.horizon_s
        ts: list[float] = []
        rps: list[float] = []
        idx = 0
        cur = w
        while cur <= end_time:                     # analyzer.py:108-125
            j = int(np.searchsorted(finish, cur, side="right"))
            ts.append(cur)
            rps.append((j - idx) / w)
            idx = j
            cur += w
        return ts, rps

    def get_sampled_metrics(self) -> dict[str, dict[str, list[float]]]:
        out: dict[str, dict[str, list[float]]] = defaultdict(dict)
        en = set(self.flat.enabled_metrics)
        if all(m in en for m in SERVER_SERIES):
            for i, sid in enumerate(self.flat.server_ids):
                for m, name in enumerate(SERVER_SERIES):
                    out[name][sid] = self.series[3 * i + m].tolist()
        else:
            for name in SERVER_SERIES:
                if name in en:
                    for sid in self.flat.server_ids:
                        out[name][sid] = []
        if EDGE_SERIES in en:
            base = 3 * self.flat.n_servers
            for j, eid in enumerate(self.flat.edge_ids):
                out[EDGE_SERIES][eid] = self.series[base + j].tolist()
        return out

    def get_metric_map(self, key: Any) -> dict[str, list[float]]:
        return self.get_sampled_metrics().get(str(getattr(key, "value", key)), {})

    def get_series(self, key: Any, entity_id: str) -> tuple[list[float], list[float]]:
        vals = self.get_metric_map(key).get(entity_id, [])
        times = (np.arange(len(vals)) * self.flat.sample_period).tolist()
        return times, vals

    # -- the reference's own analyzer, fed through its duck-typed seam ---------
    def holders(self) -> dict[str, Any]:
        sampled = self.get_sampled_metrics()
        clock = [SimpleNamespace(start=float(a), finish=float(b)) for a, b in self.clocks]
        servers = [
            SimpleNamespace(server_config=SimpleNamespace(id=sid),
                            enabled_metrics={_Name(n): sampled[n][sid] for n in SERVER_SERIES if sid in sampled.get(n, {})})
            for sid in self.flat.server_ids]
        edges = [
            SimpleNamespace(edge_config=SimpleNamespace(id=eid),
                            enabled_metrics=({_Name(EDGE_SERIES): sampled[EDGE_SERIES][eid]}
                                             if eid in sampled.get(EDGE_SERIES, {}) else {}))
            for eid in self.flat.edge_ids]
        settings = SimpleNamespace(total_simulation_time=self.flat.horizon_s,
                                   sample_period_s=self.flat.sample_period)
        return {"client": SimpleNamespace(rqs_clock=clock), "servers": servers, "edges": edges,
                "settings": settings}

    def to_reference_analyzer(self):
        """``asyncflow.metrics.analyzer.ResultsAnalyzer`` over these results (needs the reference package)."""
        from asyncflow.metrics.analyzer import ResultsAnalyzer  # noqa: PLC0415
        return ResultsAnalyzer(**self.holders())


def series_bands(replicas: list[ReplicaResults], key: Any, entity_id: str) -> dict[str, np.ndarray]:
    """Per-tick mean / min / max of one sampled series across fully traced replicas.

    ``key`` / ``entity_id`` are what ``ResultsAnalyzer.get_series`` takes (``metrics/analyzer.py:239-262``);
    returns ``{"t", "mean", "min", "max", "n"}`` -- the band a dashboard draws around one replica's curve.
    """
    cols = [np.asarray(r.get_series(key, entity_id)[1], dtype=np.float64) for r in replicas]
    cols = [c for c in cols if c.size]
    if not cols:
        z = np.zeros(0)
        return {"t": z, "mean": z, "min": z, "max": z, "n": 0}
    m = min(c.size for c in cols)
    a = np.stack([c[:m] for c in cols])
    return {"t": np.arange(m) * replicas[0].flat.sample_period, "mean": a.mean(axis=0), "min": a.min(axis=0),
            "max": a.max(axis=0), "n": len(cols)}


class SweepResults:
    """Per-replica summaries of a sweep (all arrays have one row per replica)."""

    def __init__(self, flat: FlatScenario, stats: np.ndarray, edge_sent: np.ndarray,
                 edge_dropped: np.ndarray, samp_sum: np.ndarray, samp_max: np.ndarray,
                 throughput: np.ndarray | None = None, histograms: np.ndarray | None = None,
                 replica_begin: int = 0) -> None:
        self.flat = flat
        self.stats = stats
        self.edge_sent = edge_sent
        self.edge_dropped = edge_dropped
        self.samp_sum = samp_sum
        self.samp_max = samp_max
        self.throughput = throughput
        self.histograms = histograms
        self.replica_begin = replica_begin
        self.rows: np.ndarray | None = None      # sweep row of each result row (set by SweepRunner)
        self.traced: list[ReplicaResults] = []    # full clocks + sampled series of the first `trace_replicas` replica ids
        self.exact: np.ndarray | None = None      # rows whose p50/p95/p99 are exact (SweepRunner.exact_percentiles)

    def __len__(self) -> int:
        return int(self.stats.shape[0])

    @property
    def completed(self) -> np.ndarray:
        return self.stats["completed"]

    @property
    def generated(self) -> np.ndarray:
        return self.stats["generated"]

    @property
    def overflowed(self) -> np.ndarray:
        return (self.stats["flags"] & (K.FLAG_EVENT_OVERFLOW | K.FLAG_REQUEST_OVERFLOW | K.FLAG_NOWQ_OVERFLOW)) != 0

    @property
    def mean_latency(self) -> np.ndarray:
        c = np.maximum(self.completed, 1)
        return np.where(self.completed > 0, self.stats["lat_sum"] / c, np.nan)

    @property
    def std_latency(self) -> np.ndarray:
        c = np.maximum(self.completed, 1).astype(np.float64)
        m = self.stats["lat_sum"] / c
        v = np.maximum(self.stats["lat_sumsq"] / c - m * m, 0.0)
        return np.where(self.completed > 0, np.sqrt(v), np.nan)

    def latency_stats(self, i: int) -> dict[str, float]:
        """Same keys as ``ResultsAnalyzer.get_latency_stats`` for replica ``i``;
        median/p95/p99 come from the log-linear histogram (bins <= 1.6 % wide)."""
        s = self.stats[i]
        if s["completed"] == 0:
            return {}
        return {
            "total_requests": float(s["completed"]), "mean": float(self.mean_latency[i]),
            "median": float(s["p50"]), "std_dev": float(self.std_latency[i]),
            "p95": float(s["p95"]), "p99": float(s["p99"]),
            "min": float(s["lat_min"]), "max": float(s["lat_max"]),
        }

    def sampled_mean(self) -> np.ndarray:
        """Mean over ticks of every sampled series, ``[n, n_series]``."""
        t = np.maximum(self.stats["n_ticks"], 1).astype(np.float64)[:, None]
        return self.samp_sum.astype(np.float64) / t

    def confidence_interval(self, what: str = "mean", select: np.ndarray | None = None,
                            level: float = 0.95) -> tuple[float, float, float]:
        """Monte-Carlo estimate over replicas of a per-replica statistic (``"mean"``, ``"p50"``,
        ``"p95"``, ``"p99"``, ``"completed"``, ``"generated"``): ``(estimate, lo, hi)`` with a normal
        ``level`` interval on the mean across the selected replicas (reference ROADMAP.md:23-27)."""
        if what == "mean":
            x = self.mean_latency
        elif what in ("p50", "p95", "p99", "completed", "generated", "lat_min", "lat_max"):
            x = self.stats[what].astype(np.float64)
        else:
            msg = f"unknown statistic {what!r}"
            raise KeyError(msg)
        if select is not None:
            x = x[select]
        x = x[np.isfinite(x)]
        if x.size == 0:
            return float("nan"), float("nan"), float("nan")
        m = float(x.mean())
        if x.size == 1:
            return m, m, m
        from statistics import NormalDist  # noqa: PLC0415
        z = NormalDist().inv_cdf(0.5 + level / 2.0)
        half = z * float(x.std(ddof=1)) / float(np.sqrt(x.size))
        return m, m - half, m + half

    def summary(self) -> dict[str, float]:
        ok = self.completed > 0
        tot_c = int(self.completed.sum())
        return {
            "replicas": float(len(self)), "completed": float(tot_c),
            "generated": float(self.generated.sum()),
            "events": float(self.stats["n_events"].sum()),
            "mean_latency": float(self.stats["lat_sum"].sum() / max(tot_c, 1)),
            "p50_mean": float(np.nanmean(self.stats["p50"][ok])) if ok.any() else float("nan"),
            "p95_mean": float(np.nanmean(self.stats["p95"][ok])) if ok.any() else float("nan"),
            "p99_mean": float(np.nanmean(self.stats["p99"][ok])) if ok.any() else float("nan"),
            "overflowed": float(self.overflowed.sum()),
        }

    def take(self, index: Any) -> "SweepResults":
        """Rows ``index`` (any numpy index) as a new SweepResults -- e.g. the inverse of a launch order."""
        g = lambda a: None if a is None else a[index]  # noqa: E731
        out = SweepResults(self.flat, self.stats[index], self.edge_sent[index], self.edge_dropped[index],
                           self.samp_sum[index], self.samp_max[index], g(self.throughput), g(self.histograms),
                           self.replica_begin)
        out.traced = self.traced
        return out

    def bands(self, key: Any, entity_id: str) -> dict[str, np.ndarray]:
        """``series_bands`` over this sweep's traced replicas (``SweepRunner(trace_replicas=k)``)."""
        return series_bands(self.traced, key, entity_id)

    @staticmethod
    def concatenate(parts: list["SweepResults"]) -> "SweepResults":
        f = parts[0]
        cat = lambda xs: None if xs[0] is None else np.concatenate(xs)  # noqa: E731
        return SweepResults(
            f.flat, np.concatenate([p.stats for p in parts]), np.concatenate([p.edge_sent for p in parts]),
            np.concatenate([p.edge_dropped for p in parts]), np.concatenate([p.samp_sum for p in parts]),
            np.concatenate([p.samp_max for p in parts]), cat([p.throughput for p in parts]),
            cat([p.histograms for p in parts]), f.replica_begin)
