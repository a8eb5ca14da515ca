"""Random scenario generator for the parity fuzz tests (CPU tier: twin vs oracle; GPU tier: engine
vs oracle).  Step durations come from a tiny grid so that different chains DO reach the same
instant (deterministic ties, DESIGN.md "tie rule"); servers are loaded enough to queue on CPU and
RAM; every distribution, both LB algorithms, chains, spikes and outages appear."""

from __future__ import annotations

import random

DUR = [0.001, 0.002, 0.003, 0.004, 0.008, 0.012]
CPU_KINDS = ["initial_parsing", "cpu_bound_operation"]
IO_KINDS = ["io_wait", "io_db", "io_cache", "io_llm", "io_task_spawn"]


def _endpoint(rng: random.Random, name: str) -> dict:
    steps = []
    for _ in range(rng.randint(1, 5)):
        r = rng.random()
        if r < 0.45:
            steps.append({"kind": rng.choice(CPU_KINDS), "step_operation": {"cpu_time": rng.choice(DUR)}})
        elif r < 0.85:
            steps.append({"kind": rng.choice(IO_KINDS), "step_operation": {"io_waiting_time": rng.choice(DUR)}})
        else:
            steps.append({"kind": "ram", "step_operation": {"necessary_ram": rng.choice([64, 100, 128, 200])}})
    if rng.random() < 0.7 and not any(s["kind"] == "ram" for s in steps):
        steps.insert(rng.randint(0, len(steps)), {"kind": "ram", "step_operation": {"necessary_ram": rng.choice([64, 128, 256])}})
    return {"endpoint_name": name, "steps": steps}


def _latency(rng: random.Random) -> dict:
    r = rng.random()
    if r < 0.5:
        return {"mean": rng.choice([0.001, 0.002, 0.003]), "distribution": "exponential"}
    if r < 0.7:
        m = rng.choice([0.002, 0.004])
        return {"mean": m, "distribution": "normal", "variance": m * rng.choice([0.2, 1.0])}
    if r < 0.8:
        return {"mean": 0.001, "distribution": "log_normal", "variance": 0.3}
    if r < 0.9:
        return {"mean": rng.choice([0.2, 0.6])}                        # poisson (integer seconds, mostly 0)
    return {"mean": 0.5, "distribution": "uniform"}


def scenario(seed: int) -> dict:
    rng = random.Random(seed)
    n_srv = rng.randint(1, 4)
    use_lb = n_srv > 1 and rng.random() < 0.8
    servers = []
    for i in range(n_srv):
        servers.append({
            "id": f"s{i}",
            "server_resources": {"cpu_cores": rng.choice([1, 1, 2, 3]), "ram_mb": rng.choice([256, 512, 1024])},
            "endpoints": [_endpoint(rng, f"/e{j}") for j in range(rng.randint(1, 3))],
        })
    edges = [{"id": "g-c", "source": "gen", "target": "cl", "latency": _latency(rng)}]
    events = []
    if use_lb:
        edges.append({"id": "c-lb", "source": "cl", "target": "lb", "latency": _latency(rng)})
        for i in range(n_srv):
            edges.append({"id": f"lb-s{i}", "source": "lb", "target": f"s{i}", "latency": _latency(rng),
                          "dropout_rate": rng.choice([0.0, 0.01, 0.05])})
            edges.append({"id": f"s{i}-c", "source": f"s{i}", "target": "cl", "latency": _latency(rng)})
    else:                                           # a chain cl -> s0 -> s1 ... -> cl
        edges.append({"id": "c-s0", "source": "cl", "target": "s0", "latency": _latency(rng)})
        for i in range(n_srv - 1):
            edges.append({"id": f"s{i}-s{i + 1}", "source": f"s{i}", "target": f"s{i + 1}", "latency": _latency(rng)})
        edges.append({"id": f"s{n_srv - 1}-c", "source": f"s{n_srv - 1}", "target": "cl", "latency": _latency(rng)})
    horizon = rng.choice([6, 8, 10])
    if rng.random() < 0.5:
        e = rng.choice(edges)["id"]
        events.append({"event_id": "sp1", "target_id": e,
                       "start": {"kind": "network_spike_start", "t_start": 1.0, "spike_s": rng.choice([0.004, 0.05])},
                       "end": {"kind": "network_spike_end", "t_end": 3.5}})
        events.append({"event_id": "sp2", "target_id": e,
                       "start": {"kind": "network_spike_start", "t_start": 2.0, "spike_s": 0.002},
                       "end": {"kind": "network_spike_end", "t_end": 5.0}})
    if use_lb and n_srv > 1 and rng.random() < 0.6:
        events.append({"event_id": "out1", "target_id": "s0", "start": {"kind": "server_down", "t_start": 2.0},
                       "end": {"kind": "server_up", "t_end": 4.0}})
    users = rng.choice([40, 120, 300])
    gen = {"id": "gen", "avg_active_users": {"mean": users}, "avg_request_per_minute_per_user": {"mean": rng.choice([60, 120])},
           "user_sampling_window": rng.choice([1, 3, 60])}
    if rng.random() < 0.3:
        gen["avg_active_users"] = {"mean": users, "distribution": "normal", "variance": users * 0.3}
    doc = {
        "rqs_input": gen,
        "topology_graph": {"nodes": {"client": {"id": "cl"}, "servers": servers}, "edges": edges},
        "sim_settings": {"total_simulation_time": horizon, "sample_period_s": rng.choice([0.01, 0.05])},
    }
    if use_lb:
        doc["topology_graph"]["nodes"]["load_balancer"] = {
            "id": "lb", "algorithms": rng.choice(["round_robin", "least_connection"]),
            "server_covered": [f"s{i}" for i in range(n_srv)]}
    if events:
        doc["events"] = events
    return doc


def sweep_columns(seed: int, payload: dict, n: int) -> dict:
    """Random per-replica overrides covering every AF_FIELD_* of the C ABI, for ``n`` replicas."""
    rng = random.Random(seed * 7919 + 1)
    topo = payload["topology_graph"]
    cols: dict[tuple, list] = {}
    if rng.random() < 0.7:
        cols[("users_mean",)] = [rng.choice([30, 80, 200, 350]) for _ in range(n)]
    if payload["rqs_input"]["avg_active_users"].get("distribution") == "normal" and rng.random() < 0.7:
        cols[("users_sigma",)] = [rng.choice([5.0, 20.0, 60.0]) for _ in range(n)]
    if rng.random() < 0.4:
        cols[("rate_per_user",)] = [rng.choice([0.5, 1.0, 2.0]) for _ in range(n)]
    for e in topo["edges"]:
        r = rng.random()
        dist = e["latency"].get("distribution") or "poisson"
        if r < 0.3 and dist != "uniform":
            cols[("edge_mean", e["id"])] = [rng.choice([0.001, 0.002, 0.005, 0.02]) if dist != "poisson"
                                            else rng.choice([0.1, 0.4]) for _ in range(n)]
        if 0.2 < r < 0.5 and dist in ("normal", "log_normal"):
            cols[("edge_sigma", e["id"])] = [rng.choice([0.0005, 0.002, 0.2]) for _ in range(n)]
        if r > 0.8:
            cols[("edge_dropout", e["id"])] = [rng.choice([0.0, 0.02, 0.2]) for _ in range(n)]
    for s in topo["nodes"]["servers"]:
        if rng.random() < 0.4:
            cols[("server_cpu_cores", s["id"])] = [rng.choice([1, 2, 4]) for _ in range(n)]
        if rng.random() < 0.4:
            cols[("server_ram_mb", s["id"])] = [rng.choice([256, 300, 2048]) for _ in range(n)]
        for j, ep in enumerate(s["endpoints"]):
            if rng.random() < 0.3:
                cols[("endpoint_ram", s["id"], j)] = [rng.choice([0, 32, 128, 256]) for _ in range(n)]
            for k, st in enumerate(ep["steps"]):
                if st["kind"] != "ram" and rng.random() < 0.25:
                    cols[("step_duration", s["id"], j, k)] = [rng.choice(DUR) for _ in range(n)]
    for ev in payload.get("events") or []:
        if "spike_s" in ev["start"] and rng.random() < 0.6:
            cols[("spike_delta", ev["event_id"])] = [rng.choice([0.001, 0.004, 0.03]) for _ in range(n)]
    if not cols:
        cols[("users_mean",)] = [rng.choice([30, 80, 200]) for _ in range(n)]
    return cols


def big_scenario(seed: int) -> dict:
    """C5-shaped random topologies: an LB over 5-12 front ends, some of which chain to a back end
    (several front ends may share one) before the reply returns to the client.  More edges than the
    variants' memo rows, more pending events than the sorted ring: the fallback paths get exercised."""
    rng = random.Random(seed * 104729 + 7)
    n_fe = rng.randint(5, 12)
    n_be = rng.randint(1, 4)
    servers, edges = [], []

    def server(sid: str) -> dict:
        return {"id": sid, "server_resources": {"cpu_cores": rng.choice([1, 2, 4]), "ram_mb": rng.choice([256, 512, 2048])},
                "endpoints": [_endpoint(rng, f"/e{j}") for j in range(rng.randint(1, 2))]}
    fes = [f"fe{i}" for i in range(n_fe)]
    bes = [f"be{i}" for i in range(n_be)]
    servers += [server(s) for s in fes + bes]
    edges.append({"id": "g-c", "source": "gen", "target": "cl", "latency": _latency(rng)})
    edges.append({"id": "c-lb", "source": "cl", "target": "lb", "latency": _latency(rng)})
    used_be = set()
    for f in fes:
        edges.append({"id": f"lb-{f}", "source": "lb", "target": f, "latency": _latency(rng),
                      "dropout_rate": rng.choice([0.0, 0.01])})
        if rng.random() < 0.5:
            b = rng.choice(bes)
            used_be.add(b)
            edges.append({"id": f"{f}-{b}", "source": f, "target": b, "latency": _latency(rng)})
        else:
            edges.append({"id": f"{f}-c", "source": f, "target": "cl", "latency": _latency(rng)})
    for b in bes:                                       # every server needs an exit, used or not
        edges.append({"id": f"{b}-c", "source": b, "target": "cl", "latency": _latency(rng)})
    events = []
    if rng.random() < 0.5:
        events.append({"event_id": "out1", "target_id": rng.choice(fes), "start": {"kind": "server_down", "t_start": 1.5},
                       "end": {"kind": "server_up", "t_end": 3.0}})
    if rng.random() < 0.5:
        events.append({"event_id": "sp1", "target_id": rng.choice(edges)["id"],
                       "start": {"kind": "network_spike_start", "t_start": 1.0, "spike_s": rng.choice([0.004, 0.05])},
                       "end": {"kind": "network_spike_end", "t_end": 2.5}})
    doc = {
        "rqs_input": {"id": "gen", "avg_active_users": {"mean": rng.choice([200, 600, 1500])},
                      "avg_request_per_minute_per_user": {"mean": rng.choice([60, 120])},
                      "user_sampling_window": rng.choice([1, 60])},
        "topology_graph": {"nodes": {"client": {"id": "cl"}, "servers": servers,
                                     "load_balancer": {"id": "lb", "algorithms": rng.choice(["round_robin", "least_connection"]),
                                                       "server_covered": fes}},
                           "edges": edges},
        "sim_settings": {"total_simulation_time": rng.choice([5, 6]), "sample_period_s": 0.05},
    }
    if events:
        doc["events"] = events
    return doc
