"""world_size-2 gloo test of the multi-GPU host logic (sharding + the single all-gather).

No GPU here: each rank simulates its shard with the CPU debugging twin (the same state
machine the kernel runs), then the product's `distributed` module gathers the summary
blocks.  What is checked: shards are disjoint and cover the sweep, results do not depend
on the world size (RNG keyed by global replica id), and the gathered summary equals the
single-process one."""

from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp
from helpers import SEED, load_scenario

from asyncflow_b200 import _capi as K
from asyncflow_b200.distributed import all_gather_summary, shard_bounds, summary_block
from asyncflow_b200.flatten import SweepSpec, flatten

N = 11


def _sweep(flat):
    return SweepSpec(flat, N, {("users_mean",): np.linspace(20, 200, N), ("edge_mean", "client-app"): np.linspace(0.001, 0.02, N)})


def _simulate(flat, spec, begin, end):
    import twin
    res = twin.run(flat, seed=SEED, replica_begin=begin, n=end - begin, sweep=spec, sweep_first=begin)
    return res["stats"], res["hist"].astype(np.int64).sum(axis=0)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat = flatten(load_scenario("c1_my_service.yml", 6))
    spec = _sweep(flat)
    b, e = shard_bounds(N, rank, world)
    stats, hist = _simulate(flat, spec, b, e)
    ints, flts = summary_block(stats, hist)
    g = all_gather_summary(ints, flts)
    q.put((rank, (b, e), stats["completed"].tolist(), g.completed, g.generated, g.replicas, g.histogram.tolist(),
           g.lat_sum, g.per_rank_completed))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_partition_the_sweep():
    for n in (1, 7, 100, 100_001):
        for world in (1, 2, 3, 8):
            cover = [shard_bounds(n, r, world) for r in range(world)]
            assert cover[0][0] == 0 and cover[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
            sizes = [e - b for b, e in cover]
            assert max(sizes) - min(sizes) <= 1


def test_two_ranks_equal_one_rank():
    flat = flatten(load_scenario("c1_my_service.yml", 6))
    spec = _sweep(flat)
    stats1, hist1 = _simulate(flat, spec, 0, N)
    ints, flts = summary_block(stats1, hist1)
    g1 = all_gather_summary(ints, flts)
    assert g1.completed == int(stats1["completed"].sum()) and g1.replicas == N

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, (b0, e0), c0, *g0), (r1, (b1, e1), c1, *g1r) = got
    assert (b0, e0, b1, e1) == (0, 6, 6, 11)
    # per-replica results do not depend on the world size
    assert c0 + c1 == stats1["completed"].tolist()
    # both ranks hold the same gathered summary, equal to the single-process one
    assert g0 == g1r
    completed, generated, replicas, histogram, lat_sum, per_rank = g0
    assert completed == g1.completed and generated == g1.generated and replicas == N
    assert histogram == g1.histogram.tolist()
    assert per_rank == [sum(c0), sum(c1)]
    assert abs(lat_sum - g1.lat_sum) < 1e-9 * abs(g1.lat_sum)


def test_merged_histogram_percentile_matches_numpy():
    rng = np.random.default_rng(5)
    lat = rng.lognormal(-3.7, 0.4, size=30000)
    idx = (lat.view(np.uint64) >> np.uint64(52 - K.AF_HIST_SUB_BITS)).astype(np.int64) - ((1023 + K.AF_HIST_MIN_EXP) << K.AF_HIST_SUB_BITS)
    hist = np.bincount(np.clip(idx, 0, K.AF_HIST_BINS - 1), minlength=K.AF_HIST_BINS)
    stats = np.zeros(1, dtype=K.STATS_DTYPE)
    stats["completed"] = len(lat)
    g = all_gather_summary(*summary_block(stats, hist))
    for q in (50, 95, 99):
        assert abs(g.percentile(q) - np.percentile(lat, q)) < 0.01 * np.percentile(lat, q)


def _sharded_worker(rank, world, port, q):
    """The product's run_sharded() itself, on a runner whose engine is the CPU twin."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from twin_engine import TwinEngine

    from asyncflow_b200 import SweepRunner
    from asyncflow_b200.distributed import run_sharded
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base = load_scenario("c1_my_service.yml", 6)
    sw = SweepRunner(base, N, {("users_mean",): np.linspace(20, 400, N)}, seed=SEED, pinned=False, balance=True, deal=world)
    sw._engine = TwinEngine()
    sw._engine.upload(sw.flat)
    res, g = run_sharded(sw)
    q.put((rank, res.rows.tolist(), res.completed.tolist(), g.completed, g.generated, g.replicas, g.per_rank_completed,
           int(np.asarray(g.histogram).sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_run_sharded_on_a_dealt_sweep_two_ranks():
    from twin_engine import TwinEngine

    from asyncflow_b200 import SweepRunner
    base = load_scenario("c1_my_service.yml", 6)
    one = SweepRunner(base, N, {("users_mean",): np.linspace(20, 400, N)}, seed=SEED, pinned=False, balance=True, deal=2)
    one._engine = TwinEngine()
    one._engine.upload(one.flat)
    full = one.run()

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, rows0, c0, *g0), (_, rows1, c1, *g1) = got
    assert sorted(rows0 + rows1) == list(range(N))
    for rows, comp in ((rows0, c0), (rows1, c1)):
        assert comp == [int(full.completed[r]) for r in rows]          # same result per row as the one-process run
    assert g0 == g1
    completed, generated, replicas, per_rank, hist_total = g0
    assert completed == int(full.completed.sum()) and generated == int(full.generated.sum()) and replicas == N
    assert per_rank == [sum(c0), sum(c1)] and hist_total == completed
    # dealing balanced the two ranks' work (the undealt contiguous split would be ~1:3)
    assert max(per_rank) / min(per_rank) < 1.35
