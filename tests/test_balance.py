"""Launch order of skewed sweeps (flatten.balanced_order, SweepRunner(balance=True)): host logic only.

The C2 sweep (users 10 -> 1000) is sorted by ascending load; the engine hands replica
ids to warps in order, so the order in which rows get ids decides the tail of the launch (one GPU) and
the balance between ranks (several)."""

from __future__ import annotations

import des_port
import numpy as np
import twin
from helpers import SEED, assert_matches_oracle, load_scenario

from asyncflow_b200 import SweepResults, SweepRunner, balanced_order, flatten
from asyncflow_b200 import _capi as K
from asyncflow_b200.distributed import shard_bounds


def test_order_is_a_heaviest_first_permutation():
    rng = np.random.default_rng(1)
    cost = rng.uniform(1, 100, 1000)
    o = balanced_order(cost)
    assert sorted(o.tolist()) == list(range(1000))
    assert (np.diff(cost[o]) <= 0).all()
    # flat or already sorted sweeps are left alone; ties keep their row order
    assert balanced_order(np.ones(7)).tolist() == list(range(7))
    assert balanced_order([5, 5, 9, 5]).tolist() == [2, 0, 1, 3]


def test_dealing_gives_every_rank_the_same_mix():
    cost = np.linspace(10, 1000, 10_000) ** 1.3          # C2-like: ascending, convex
    for world in (2, 4, 8):
        o = balanced_order(cost, deal=world)
        assert sorted(o.tolist()) == list(range(cost.size))
        share = []
        for r in range(world):
            b, e = shard_bounds(cost.size, r, world)
            share.append(cost[o[b:e]].sum())
            assert (np.diff(cost[o[b:e]]) <= 0).all()     # heaviest first inside every shard
        assert max(share) / min(share) < 1.01
        # the contiguous split of the unsorted sweep is what this repairs
        plain = [cost[slice(*shard_bounds(cost.size, r, world))].sum() for r in range(world)]
        assert max(plain) / min(plain) > 2.0


def test_ascending_sweep_tail_model():
    """List scheduling on W slots in id order: ascending order ends with the heaviest replica alone."""
    cost = np.linspace(10, 1000, 10_000)
    W = 4736                                              # 148 SMs x 32 warps

    def makespan(c):
        import heapq
        slots = [0.0] * W
        for x in c:
            heapq.heappush(slots, heapq.heappop(slots) + x)
        return max(slots)
    asc, lpt = makespan(cost), makespan(cost[balanced_order(cost)])
    ideal = cost.sum() / W
    assert lpt < 1.10 * max(ideal, cost.max())         # 1124 vs 1066 (ascending: 1593)
    assert asc > 1.3 * lpt


def test_balanced_runner_keeps_rows_and_names_their_replica_ids():
    base = load_scenario("c1_my_service.yml", 6)
    users = [30.0, 200.0, 10.0, 120.0, 60.0]
    rtt = [0.001, 0.002, 0.003, 0.004, 0.005]
    sw = SweepRunner(base, 5, {("users_mean",): users, ("edge_mean", "client-app"): rtt}, pinned=False, balance=True)
    assert sw.order.tolist() == [1, 3, 4, 0, 2]
    assert sw.replica_ids.tolist() == [3, 0, 4, 1, 2]
    assert sw.rows_of(1, 3).tolist() == [3, 4]
    # what the engine is given: position p carries row order[p]
    np.testing.assert_array_equal(sw._run_spec.values[:, 0], np.array(users)[sw.order])
    np.testing.assert_array_equal(sw._run_spec.values[:, 1], np.array(rtt)[sw.order])
    # the engine's state machine on that table (CPU twin) == the oracle on each ROW's payload with ITS replica id
    flat = sw.flat
    r = twin.run(flat, seed=SEED, replica_begin=0, n=5, sweep=sw._run_spec, trace=5, clock_cap=20000)
    for row in range(5):
        rid = int(sw.replica_ids[row])
        p = sw.payload_for(row)
        assert p["rqs_input"]["avg_active_users"]["mean"] == users[row]
        o = des_port.simulate(p, seed=SEED, replica=rid)
        k = int(r["stats"][rid]["completed"])
        assert_matches_oracle(o, flat, stats=r["stats"][rid], clocks=r["trace_clocks"][rid, :k],
                              sent=r["sent"][rid], dropped=r["dropped"][rid])
        rr = sw.replica_runner(row)
        assert (rr.seed, rr.replica) == (sw.seed, rid)
    # un-permuting the collected arrays
    res = SweepResults(flat, r["stats"], r["sent"], r["dropped"], r["samp_sum"], r["samp_max"])
    back = res.take(sw.replica_ids)
    assert back.generated.tolist() == [int(r["stats"][sw.replica_ids[i]]["generated"]) for i in range(5)]
    assert np.argmax(back.generated) == 1 and np.argmin(back.generated) == 2      # users 200 / users 10


def test_unbalanced_runner_is_unchanged():
    base = load_scenario("c1_my_service.yml", 6)
    sw = SweepRunner(base, 3, {("users_mean",): [10, 20, 30]}, pinned=False)
    assert sw.order is None and sw._run_spec is sw.spec
    assert sw.replica_ids.tolist() == [0, 1, 2] and sw.rows_of(1, 3).tolist() == [1, 2]
    flat_only = SweepRunner(base, 3, {("edge_mean", "client-app"): [0.001, 0.002, 0.003]}, pinned=False, balance=True)
    assert flat_only.order is None                        # nothing that changes the load is swept
    assert K.FIELDS["users_mean"] == 0
