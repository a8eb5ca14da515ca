"""A skewed sweep (BASELINE.json configs[1]: one server, users 10 -> 1000) with drill-down.

    python examples/sweep_users_drilldown.py [n_replicas]

* ``balance=True`` launches the saturated rows first (the sweep is sorted by ascending load);
* ``trace_replicas=8`` keeps full traces of the first eight replica ids (the heaviest rows) and
  ``bands`` gives the per-tick mean/min/max of a sampled series across them;
* ``replica_runner(row)`` replays one sweep point with every (start, finish) clock and sampled
  series, ``payload_for(row)`` is that point as a plain payload for the reference's own runner.
"""
import sys
from pathlib import Path

import numpy as np
import yaml

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from asyncflow_b200 import SweepRunner  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
payload = yaml.safe_load((ROOT / "tests" / "scenarios" / "c1_my_service.yml").read_text())
payload["sim_settings"]["total_simulation_time"] = 60
users = np.linspace(10, 1000, n)
sweep = SweepRunner(payload, n, {("users_mean",): users}, seed=11, balance=True, trace_replicas=8)
res = sweep.run()
ms_total, ms_sim = sweep.last_ms
print(f"{n} replicas, {int(res.completed.sum()):,} completions, kernel {ms_sim:.0f} ms "
      f"({res.completed.sum() / ms_sim * 1e3:.3g} completions/s), overflowed: {int(res.overflowed.sum())}")
for lo, hi in ((10, 100), (250, 350), (900, 1000)):
    sel = (users >= lo) & (users <= hi)
    m, a, b = res.confidence_interval("p95", sel)
    print(f"users {lo:4d}-{hi:4d}: p95 {m * 1e3:8.1f} ms  (95% CI {a * 1e3:.1f}-{b * 1e3:.1f} over {sel.sum()} replicas)")

band = res.bands("ready_queue_len", "app-1")
print(f"ready queue of app-1 over the {band['n']} heaviest rows: mean of means {band['mean'].mean():.1f}, peak {band['max'].max():.0f}")

row = int(np.searchsorted(users, 400.0))
one = sweep.replica_runner(row).run()                   # the same random numbers row `row` had in the sweep
print(f"row {row} (users {users[row]:.0f}) replayed: {one.format_latency_stats()}")
assert int(res.completed[row]) == one.clocks.shape[0]
print("as a reference payload:", sweep.payload_for(row)["rqs_input"])
