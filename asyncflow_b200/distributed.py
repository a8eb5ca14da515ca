"""Multi-GPU: replicas shard by range, one collective at the end.

Replicas are independent and AF-RNG is keyed by the GLOBAL replica id, so a sweep
over G GPUs is G disjoint replica ranges with zero traffic during simulation and
results that do not depend on G.  The only exchange is one all-gather of each
rank's *summary block* (the device-reduced latency histogram plus totals; a few
KB) over NCCL/NVLink -- BASELINE.json north_star, SURVEY.md 8e.  The reference
has nothing to mirror here (single process, single thread).

One process per GPU (``torchrun``); ``torch.distributed`` is plumbing only.  The
same code runs on the ``gloo`` backend for the CPU-tier tests.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _capi as K

N_TOTALS = 8     # completed, generated, events, replicas, overflowed, ticks, reserved x2


def shard_bounds(n_replicas: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous replica range of ``rank``: sizes differ by at most one."""
    base, extra = divmod(int(n_replicas), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def summary_block(stats: np.ndarray, reduced_hist: np.ndarray | None) -> tuple[np.ndarray, np.ndarray]:
    """(int64 block, float64 block) a rank contributes to the final all-gather."""
    ints = np.zeros(K.AF_HIST_BINS + N_TOTALS, dtype=np.int64)
    if reduced_hist is not None:
        ints[: K.AF_HIST_BINS] = reduced_hist.astype(np.int64)
    t = ints[K.AF_HIST_BINS:]
    t[0] = int(stats["completed"].sum())
    t[1] = int(stats["generated"].sum())
    t[2] = int(stats["n_events"].sum())
    t[3] = int(stats.shape[0])
    t[4] = int(((stats["flags"] & (K.FLAG_EVENT_OVERFLOW | K.FLAG_REQUEST_OVERFLOW | K.FLAG_NOWQ_OVERFLOW)) != 0).sum())
    t[5] = int(stats["n_ticks"].sum())
    flts = np.array([float(stats["lat_sum"].sum()), float(stats["lat_sumsq"].sum()),
                     float(stats["lat_min"][stats["completed"] > 0].min()) if (stats["completed"] > 0).any() else np.inf,
                     float(stats["lat_max"].max()) if stats.shape[0] else 0.0], dtype=np.float64)
    return ints, flts


@dataclass
class GlobalSummary:
    histogram: np.ndarray      # [AF_HIST_BINS] summed over every rank
    completed: int
    generated: int
    events: int
    replicas: int
    overflowed: int
    lat_sum: float
    lat_min: float
    lat_max: float
    per_rank_completed: list[int]

    @property
    def mean_latency(self) -> float:
        return self.lat_sum / max(self.completed, 1)

    def percentile(self, q: float) -> float:
        """numpy-'linear' percentile of ALL completions, read off the merged histogram."""
        n = int(self.histogram.sum())
        if n == 0:
            return float("nan")
        cum = np.cumsum(self.histogram)
        base = (1023 + K.AF_HIST_MIN_EXP) << K.AF_HIST_SUB_BITS

        def order_stat(rank: int) -> float:
            b = int(np.searchsorted(cum, rank, side="right"))
            before = int(cum[b - 1]) if b else 0
            lo, hi = (np.array([(b + base) << (52 - K.AF_HIST_SUB_BITS),
                                (b + 1 + base) << (52 - K.AF_HIST_SUB_BITS)], dtype=np.uint64).view(np.float64))
            frac = ((rank - before) + 0.5) / float(self.histogram[b])
            return float(lo + frac * (hi - lo))

        pos = q / 100.0 * (n - 1)
        lo = int(np.floor(pos))
        frac = pos - lo
        a = order_stat(lo)
        return a if frac == 0.0 or lo + 1 >= n else a + frac * (order_stat(lo + 1) - a)


def all_gather_summary(ints: np.ndarray, flts: np.ndarray, device: str | None = None) -> GlobalSummary:
    """One all-gather of the summary blocks (NCCL on GPUs, gloo on CPU); works without
    an initialised process group (world = 1)."""
    import torch  # noqa: PLC0415
    import torch.distributed as dist  # noqa: PLC0415

    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        all_i, all_f = ints[None, :], flts[None, :]
    else:
        dev = device or ("cuda" if dist.get_backend() == "nccl" else "cpu")
        ti = torch.from_numpy(ints).to(dev)
        tf = torch.from_numpy(flts).to(dev)
        gi = [torch.empty_like(ti) for _ in range(world)]
        gf = [torch.empty_like(tf) for _ in range(world)]
        dist.all_gather(gi, ti)
        dist.all_gather(gf, tf)
        all_i = torch.stack(gi).cpu().numpy()
        all_f = torch.stack(gf).cpu().numpy()
    tot = all_i[:, K.AF_HIST_BINS:]
    return GlobalSummary(
        histogram=all_i[:, : K.AF_HIST_BINS].sum(axis=0).astype(np.int64),
        completed=int(tot[:, 0].sum()), generated=int(tot[:, 1].sum()), events=int(tot[:, 2].sum()),
        replicas=int(tot[:, 3].sum()), overflowed=int(tot[:, 4].sum()),
        lat_sum=float(all_f[:, 0].sum()), lat_min=float(all_f[:, 2].min()), lat_max=float(all_f[:, 3].max()),
        per_rank_completed=[int(x) for x in tot[:, 0]])


def run_sharded(runner, *, with_histogram: bool = True):
    """Run ``runner`` (a :class:`~asyncflow_b200.runner.SweepRunner` built for the WHOLE sweep)
    on this rank's replica range and all-gather the summaries.

    One process per GPU (``torchrun``); call ``torch.distributed.init_process_group("nccl")``
    first.  Returns ``(local SweepResults, GlobalSummary)``; the per-replica rows of other
    ranks stay on those ranks -- only the summary block crosses NVLink.
    """
    import torch.distributed as dist  # noqa: PLC0415

    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    begin, end = shard_bounds(runner.n_replicas, rank, world)
    if end <= begin:
        # more ranks than replicas: nothing to launch here, but the collective must still complete
        from . import _capi as K  # noqa: PLC0415
        res = None
        ints, flts = summary_block(np.zeros(0, dtype=K.STATS_DTYPE), None)
        return res, all_gather_summary(ints, flts)
    res = runner.run(begin, end)
    hist = runner.engine().reduced_histogram() if (with_histogram and runner.histogram) else None
    ints, flts = summary_block(res.stats, hist)
    return res, all_gather_summary(ints, flts)
