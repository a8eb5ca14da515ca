"""Exploration helper (not the contract bench): time the sim kernel on one GPU."""
import argparse, sys, time
from pathlib import Path
import numpy as np, yaml
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from asyncflow_b200 import SweepRunner, flatten

ap = argparse.ArgumentParser()
ap.add_argument("--scenario", default="c3_lb_two_servers.yml")
ap.add_argument("--replicas", type=int, default=20000)
ap.add_argument("--horizon", type=int, default=30)
ap.add_argument("--wpb", type=int, default=0)
ap.add_argument("--bps", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--sweep", default="rtt")
ap.add_argument("--no-metrics", action="store_true")
ap.add_argument("--users", type=float, default=0.0)
ap.add_argument("--balance", action="store_true", help="launch heaviest rows first (SweepRunner(balance=True))")
ap.add_argument("--mode", default="", help="auto | warp | lane (af_engine_set_mode); default: the library's")
ap.add_argument("--event-capacity", type=int, default=0)
ap.add_argument("--request-capacity", type=int, default=None)
a = ap.parse_args()
d = yaml.safe_load((ROOT / "tests" / "scenarios" / a.scenario).read_text())
d["sim_settings"]["total_simulation_time"] = a.horizon
if a.no_metrics:
    d["sim_settings"]["enabled_sample_metrics"] = []
if a.users > 0:
    d["rqs_input"]["avg_active_users"]["mean"] = a.users
flat = flatten(d)
n = a.replicas
sweep = None
if a.sweep == "rtt":
    rtt = np.linspace(0.001, 0.050, n)
    sweep = {("edge_mean", e): rtt for e in flat.edge_ids}
elif a.sweep == "users":
    sweep = {("users_mean",): np.linspace(10, 1000, n)}
sw = SweepRunner(flat, n, sweep, warps_per_block=a.wpb, blocks_per_sm=a.bps, balance=a.balance,
                 event_capacity=a.event_capacity, request_capacity=a.request_capacity)
if a.mode:
    sw.engine().set_mode(a.mode)
for i in range(a.reps):
    t = time.time(); res = sw.run(); wall = time.time() - t
    ms_total, ms_sim = sw.last_ms
    s = res.summary()
    print(f"run{i}: wall {wall*1e3:.1f} ms, device total {ms_total:.2f} ms, sim {ms_sim:.2f} ms, "
          f"completed {s['completed']:.3e}, events {s['events']:.3e}, "
          f"{s['completed']/ms_sim*1e3:.3e} compl/s, {s['events']/ms_sim*1e3:.3e} events/s, overflow {s['overflowed']}, "
          f"peak_ev {res.stats['peak_events'].max()} peak_rq {res.stats['peak_requests'].max()} mean_lat {s['mean_latency']:.5f}", flush=True)
try:
    print("passes:", sw.engine().last_run_passes(), flush=True)
except AttributeError:
    pass
