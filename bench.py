#!/usr/bin/env python
"""bench.py -- simulated request-completions/s of the replica engine (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], the one the north-star target is quoted on):
client -> LB -> {srv-1, srv-2} (README dashboard example), 100 000 replicas per GPU,
every edge's latency swept over RTT 1-50 ms (mean) x jitter 10-50 % (normal, sigma =
jitter * mean), fixed seed.  A *step* is one full pass of the hot path over that
batch: every replica simulated from t=0 to the horizon.

* value  -- whole-job completions/s, sweep rows already resident in HBM, device-timed
            (CUDA events on the engine's stream; max over ranks).
* e2e    -- the same through SweepRunner.run(): pinned-host sweep rows H2D, simulation,
            per-replica statistics / edge counters / sampled aggregates D2H.
* N > 1  -- replicas shard by range, no traffic during simulation, one NCCL all-gather
            of each rank's summary block (reduced latency histogram + totals) per step.

`--impl reference` times the reference's CPU path (oracle/des_port.py: the actor
generators on a simpy-4.1.1-compatible heap, restated because simpy is not installable
here) on all host cores, on a bounded sample of the same workload.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import yaml

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SEED = 0xA5F10
BYTES_PER_COMPLETION = 680.0     # 7 timed events x 96 B + 8 B histogram RMW (LB topology)
METRIC = "simulated request-completions/sec"
UNIT = "completions/s"


# --------------------------------------------------------------------------- workload
def workload(n_replicas: int, horizon: int):
    payload = yaml.safe_load((ROOT / "tests" / "scenarios" / "c3_lb_two_servers.yml").read_text())
    payload["sim_settings"]["total_simulation_time"] = horizon
    for e in payload["topology_graph"]["edges"]:          # "+ jitter": normal latency, sigma swept
        e["latency"] = {"mean": e["latency"]["mean"], "distribution": "normal",
                        "variance": 0.3 * e["latency"]["mean"]}
    return payload


def sweep_rows(replica_ids: np.ndarray, total: int):
    """RTT x jitter grid, a pure function of the GLOBAL replica id."""
    n_j = 100
    n_r = max(total // n_j, 1)
    rtt = 0.001 + (0.050 - 0.001) * ((replica_ids // n_j) % n_r) / max(n_r - 1, 1)
    jit = 0.1 + 0.4 * (replica_ids % n_j) / (n_j - 1)
    return rtt, jit * rtt


def edge_ids(payload) -> list[str]:
    return [e["id"] for e in payload["topology_graph"]["edges"]]


# --------------------------------------------------------------------------- CPU path
def _cpu_one(args):
    payload, seed, replica, mean, sigma = args
    sys.path[:0] = [str(ROOT / "oracle"), str(ROOT / "oracle" / "simpy_shim")]
    import des_port
    for e in payload["topology_graph"]["edges"]:
        e["latency"]["mean"] = float(mean)
        e["latency"]["variance"] = float(sigma)
    r = des_port.simulate(payload, seed=seed, replica=replica)
    return r["completed"], r["heap_events"]


_POOL = None


def _noop(_):
    return 0


def cpu_pool(cores: int):
    """A warm process pool (fork + imports are NOT part of what gets timed)."""
    global _POOL
    if _POOL is None and cores > 1:
        import multiprocessing as mp
        _POOL = mp.get_context("fork").Pool(cores)
        _POOL.map(_noop, range(cores * 4))
    return _POOL


def cpu_pool_close() -> None:
    global _POOL
    if _POOL is not None:
        _POOL.close()
        _POOL.join()
        _POOL = None


def cpu_path(payload, replica_ids, total, cores: int):
    """Simulate `replica_ids` with the reference's CPU path on `cores` processes."""
    rtt, sig = sweep_rows(np.asarray(replica_ids), total)
    jobs = [(payload, SEED, int(r), m, s) for r, m, s in zip(replica_ids, rtt, sig)]
    pool = cpu_pool(cores)
    t0 = time.perf_counter()
    out = pool.map(_cpu_one, jobs, chunksize=1) if pool is not None else [_cpu_one(j) for j in jobs]
    dt = time.perf_counter() - t0
    return sum(c for c, _ in out), sum(h for _, h in out), dt


def spaced(total: int, k: int) -> np.ndarray:
    return np.unique(np.linspace(0, total - 1, k).astype(np.int64))


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.rows: list[list[str]] = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self) -> None:
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def latency_delta_vs_reference(device: int):
    """BASELINE.json's second half: p50/p95/p99 latency delta vs the SimPy reference.  The reference
    side is tests/golden/reference_numpy_stats.json (the unmodified reference on its own numpy RNG,
    96 replicas of C3 at the reference's parameters, 60 s horizon: oracle/make_reference_stats.py);
    the engine side is a 4096-replica run of the same scenario, outside the timed region."""
    from asyncflow_b200 import SweepRunner, flatten
    fx = ROOT / "tests" / "golden" / "reference_numpy_stats.json"
    if not fx.exists():
        return None
    ref = json.loads(fx.read_text())["c3_lb_two_servers.yml"]
    payload = yaml.safe_load((ROOT / "tests" / "scenarios" / "c3_lb_two_servers.yml").read_text())
    payload["sim_settings"]["total_simulation_time"] = ref["horizon"]
    sw = SweepRunner(flatten(payload), 4096, seed=SEED + 1, device=device)
    st = sw.run().stats
    sw.close()
    mine = {"mean": float((st["lat_sum"] / st["completed"]).mean()), "median": float(st["p50"].mean()),
            "p95": float(st["p95"].mean()), "p99": float(st["p99"].mean())}
    return {"scenario": "c3_lb_two_servers.yml (reference parameters, horizon %d s)" % ref["horizon"],
            "reference": "unmodified AsyncFlow actors, numpy PCG64, %d replicas" % ref["replicas"],
            "engine_replicas": 4096,
            "delta_pct": {k: 100.0 * (mine[k] - ref["mean"][k]) / ref["mean"][k] for k in mine},
            "reference_s": {k: ref["mean"][k] for k in mine}, "engine_s": mine}


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except (KeyError, ValueError):
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(config_key: str):
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            entry = d.get(config_key)
            return None if entry is None else entry.get("dram_bytes_per_launch")
        except ValueError:
            return None
    return None


# --------------------------------------------------------------------------- arms
def run_reference(a) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    total = a.replicas * a.gpus
    payload = workload(a.replicas, a.horizon)
    per_step = 2 * cores                                  # ~2 replicas per core and step (~15 s)
    ids = spaced(total, per_step * (a.steps + a.warmup))
    chunks = [ids[i::(a.steps + a.warmup)] for i in range(a.steps + a.warmup)]
    for c in chunks[: a.warmup]:
        cpu_path(payload, c, total, cores)
    comp = ev = 0
    dt = 0.0
    for c in chunks[a.warmup:]:
        n, h, t = cpu_path(payload, c, total, cores)
        comp += n; ev += h; dt += t
    value = comp / dt
    sample = (f"{len(chunks[0])} replicas/step spaced over the sweep, horizon {a.horizon}s, "
              f"{cores} processes (multiprocessing), oracle/des_port.py on oracle/simpy_shim")
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(a),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "heap_events_per_s": ev / dt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def config_dict(a) -> dict:
    return {"workload": "configs[2]: client->LB->{srv-1,srv-2} (README dashboard example), "
                        f"{a.replicas} replicas/GPU sweeping edge RTT 1-50 ms x jitter 10-50 % (normal)",
            "replicas_per_gpu": a.replicas, "horizon_s": a.horizon,
            "horizon_note": "reference YAML horizon is 600 s; the metric is a rate, the horizon only scales step length",
            "seed": hex(SEED), "parallelism": f"replica-range x{a.gpus}",
            "l2": "working set per step (819 MB of per-replica histograms at 100k replicas) exceeds the 126 MB L2; "
                  "a 256 MB buffer is also overwritten between timed steps"}


def run_ours(a) -> None:
    import torch
    import torch.distributed as dist

    from asyncflow_b200 import SweepRunner, flatten
    from asyncflow_b200._capi import STATS_DTYPE as res_dtype
    from asyncflow_b200.distributed import all_gather_summary, summary_block

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device (asyncflow_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    total = a.replicas * world
    begin = rank * a.replicas
    payload = workload(a.replicas, a.horizon)
    flat = flatten(payload)
    ids = np.arange(begin, begin + a.replicas, dtype=np.int64)
    rtt, sig = sweep_rows(ids, total)
    cols = {}
    for e in flat.edge_ids:
        cols[("edge_mean", e)] = rtt
        cols[("edge_sigma", e)] = sig
    # one SweepRunner per rank holding this rank's rows; replica ids stay global
    sw = SweepRunner(flat, a.replicas, cols, seed=SEED, device=local, histogram=True, throughput=False)
    eng = sw.engine()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def launch():
        eng.configure(request_capacity=sw.request_capacity, event_capacity=sw.event_capacity,
                      histogram=True, throughput=False)
        eng.run(SEED, begin, begin + a.replicas)

    empty_stats = np.zeros(0, dtype=res_dtype)

    def summarise(res_stats=None):
        """This rank's summary block (device-reduced histogram + totals), all-gathered over
        NCCL when world > 1 -- the sweep's only collective."""
        ints, flts = summary_block(empty_stats if res_stats is None else res_stats, eng.reduced_histogram())
        return all_gather_summary(ints, flts, device="cuda" if world > 1 else None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step():
        launch()
        eng.sync()
        ms_total, ms_sim = eng.last_run_ms()
        summarise()
        return ms_total, ms_sim

    def e2e_step():
        eng.upload_sweep(sw.spec, begin, row_first=0, row_count=a.replicas)
        launch()
        res = sw.collect(begin)
        res.global_summary = summarise(res.stats)
        return res

    # sweep rows resident for the `value` steps
    eng.upload_sweep(sw.spec, begin, row_first=0, row_count=a.replicas)
    for _ in range(a.warmup):
        resident_step()
    sampler = ClockSampler(local) if rank == 0 else None

    # ---- value: K steps, inputs resident, device-timed ---------------------------
    launches0 = eng.launch_count
    dev_ms = sim_ms = 0.0
    barrier()
    w0 = time.perf_counter()
    for _ in range(a.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        mt, ms = resident_step()
        dev_ms += mt; sim_ms += ms
    barrier()
    wall_resident = time.perf_counter() - w0
    launches = eng.launch_count - launches0

    # ---- e2e: K steps through the public API, host buffers ------------------------
    res = e2e_step()                                    # untimed: allocates the pinned result buffers
    barrier()
    w0 = time.perf_counter()
    for _ in range(a.steps):
        res = e2e_step()
    barrier()
    wall_e2e = time.perf_counter() - w0
    clocks = sampler.stop() if sampler else None

    st = res.stats
    mine = np.array([float(st["completed"].sum()), float(st["n_events"].sum()), dev_ms, sim_ms, wall_e2e,
                     wall_resident, float(res.overflowed.sum())])
    if world > 1:
        t = torch.tensor(mine, device="cuda", dtype=torch.float64)
        tot = t.clone(); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        completed, events, overflow = tot[0].item(), tot[1].item(), tot[6].item()
        dev_ms, sim_ms, wall_e2e, wall_resident = mx[2].item(), mx[3].item(), mx[4].item(), mx[5].item()
    else:
        completed, events, overflow = mine[0], mine[1], mine[6]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = completed * a.steps / (dev_ms / 1e3)
    e2e_value = completed * a.steps / wall_e2e
    peak, peak_src = hbm_peak()
    per_launch_completions = float(st["completed"].sum())
    achieved = per_launch_completions * BYTES_PER_COMPLETION / (sim_ms / a.steps / 1e3) / 1e9
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_dict(a),
        "events_per_s": events * a.steps / (dev_ms / 1e3),
        "wall_ms_per_step_resident": wall_resident / a.steps * 1e3,
        "replicas_overflowed": overflow,
        "latency_all_replicas": {"mean_s": res.global_summary.mean_latency, "p50_s": res.global_summary.percentile(50),
                                 "p95_s": res.global_summary.percentile(95), "p99_s": res.global_summary.percentile(99),
                                 "source": "merged (all-gathered) histogram"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(sw.spec.values.nbytes),
                "d2h_bytes_per_step": int(sw.d2h_bytes + (2048 * 8)), "ms_per_step": wall_e2e / a.steps * 1e3},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "af_sim_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src,
                     "algorithmic_bytes_per_completion": BYTES_PER_COMPLETION,
                     "traffic": profiled_traffic(f"c3_r{a.replicas}_t{a.horizon}"),
                     "note": "latency/issue-bound state machine: HBM fraction is not the limiter (DESIGN.md 'Roofline')"},
    }
    out["latency_delta_vs_reference"] = latency_delta_vs_reference(local)
    if world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        k = 3 * cores                                     # ~3 replicas per core: 10-30 s of CPU work
        cpu_path(payload, spaced(total, cores), total, cores)   # untimed: page in the interpreter state
        n, h, dt = cpu_path(payload, spaced(total, k), total, cores)
        out["cpu_baseline"] = {
            "value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{k} replicas spaced over the sweep, horizon {a.horizon}s, {cores} processes, "
                      f"oracle/des_port.py on oracle/simpy_shim ({dt:.1f} s)",
            "heap_events_per_s": h / dt}
    emit(out)
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def claim_stdout() -> None:
    """Keep stdout for the ONE JSON line: native libraries write there too (NCCL prints its version
    banner with printf at communicator creation), so fd 1 is pointed at stderr for the run and the
    result goes to the saved descriptor."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(obj: dict) -> None:
    sys.stdout.flush()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, (json.dumps(obj) + "\n").encode())


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--replicas", type=int, default=100_000, help="replicas per GPU")
    ap.add_argument("--horizon", type=int, default=60)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 0)
    claim_stdout()
    try:
        if a.impl == "reference":
            run_reference(a)
        else:
            run_ours(a)
    finally:
        cpu_pool_close()


if __name__ == "__main__":
    main()
