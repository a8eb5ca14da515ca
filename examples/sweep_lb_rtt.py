"""Monte-Carlo sweep: the README LB example over edge RTT x jitter (BASELINE.json configs[2]).

    python examples/sweep_lb_rtt.py [n_replicas]
"""
import sys
from pathlib import Path

import numpy as np
import yaml

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from asyncflow_b200 import SweepRunner, flatten  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000
payload = yaml.safe_load((ROOT / "tests" / "scenarios" / "c3_lb_two_servers.yml").read_text())
payload["sim_settings"]["total_simulation_time"] = 60
for e in payload["topology_graph"]["edges"]:
    e["latency"] = {"mean": e["latency"]["mean"], "distribution": "normal", "variance": 0.001}
flat = flatten(payload)
rtt = np.repeat(np.linspace(0.001, 0.050, n // 100), 100)[:n]
jitter = np.tile(np.linspace(0.1, 0.5, 100), n // 100 + 1)[:n]
cols = {}
for e in flat.edge_ids:
    cols[("edge_mean", e)] = rtt
    cols[("edge_sigma", e)] = jitter * rtt
sweep = SweepRunner(flat, n, cols, seed=7)
res = sweep.run()
ms_total, ms_sim = sweep.last_ms
print(f"{n} replicas, {int(res.completed.sum()):,} completions in {ms_sim:.0f} ms of kernel time "
      f"({res.completed.sum() / ms_sim * 1e3:.3g} completions/s)")
for lo, hi in ((0.001, 0.010), (0.020, 0.030), (0.040, 0.050)):
    sel = (rtt >= lo) & (rtt <= hi)
    ci = res.confidence_interval("p95", sel)
    print(f"RTT {lo * 1e3:.0f}-{hi * 1e3:.0f} ms: p95 latency {ci[0] * 1e3:.1f} ms  (95% CI {ci[1] * 1e3:.1f}-{ci[2] * 1e3:.1f} ms over {sel.sum()} replicas)")
