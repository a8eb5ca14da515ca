"""The OUT seam: engine results fed to the reference's REAL ResultsAnalyzer.

Needs /root/reference (build container, no GPU), so the engine results come from the CPU
debugging twin; the object under test is the product's `ReplicaResults` (holders, getters,
`to_reference_analyzer`).  On the GPU box the same class is exercised with real engine output
by tests/test_gpu_parity.py::test_runner_api_mirrors_the_reference."""

from __future__ import annotations

import numpy as np
import pytest
import twin
from helpers import SEED, load_scenario

import ref_harness
from asyncflow_b200.flatten import flatten
from asyncflow_b200.results import ReplicaResults

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_harness.reference_available(), reason="/root/reference not on this box")]


def twin_results(payload, replica) -> ReplicaResults:
    flat = flatten(payload)
    r = twin.run(flat, seed=SEED, replica_begin=replica, n=1, trace=1, clock_cap=100000)
    st = r["stats"][0]
    n, nt = int(st["completed"]), int(st["n_ticks"])
    return ReplicaResults(flat=flat, clocks=r["trace_clocks"][0, :n].copy(), series=r["trace_series"][0][:, :nt].copy(),
                          generated=int(st["generated"]), edge_sent=dict(zip(flat.edge_ids, map(int, r["sent"][0]))),
                          edge_dropped=dict(zip(flat.edge_ids, map(int, r["dropped"][0]))),
                          n_events=int(st["n_events"]), flags=int(st["flags"]))


@pytest.mark.parametrize("name,horizon", [("c1_my_service.yml", 15), ("ev_spikes_outages.yml", None), ("mixed_lc.yml", None)])
def test_reference_analyzer_on_engine_results_equals_reference_run(name, horizon):
    payload = load_scenario(name, horizon)
    ref = ref_harness.run_reference(payload, seed=SEED, replica=4)
    mine = twin_results(payload, 4)
    ra, rb = ref["analyzer"], mine.to_reference_analyzer()       # both are asyncflow ResultsAnalyzer
    assert type(ra) is type(rb)
    sa, sb = ra.get_latency_stats(), rb.get_latency_stats()
    assert {k.value: v for k, v in sa.items()} == {k.value: v for k, v in sb.items()}
    assert ra.get_throughput_series() == rb.get_throughput_series()
    assert ra.get_throughput_series(window_s=2.5) == rb.get_throughput_series(window_s=2.5)
    ma, mb = ra.get_sampled_metrics(), rb.get_sampled_metrics()
    assert set(ma) == set(mb)
    for metric in ma:
        assert set(ma[metric]) == set(mb[metric])
        for ent in ma[metric]:
            assert list(ma[metric][ent]) == list(mb[metric][ent]), (metric, ent)
    assert ra.list_server_ids() == rb.list_server_ids()
    # and the product's own getters agree with the reference analyzer's
    own = mine.get_latency_stats()
    assert own == {k.value: v for k, v in sa.items()}
    assert mine.get_throughput_series() == ra.get_throughput_series()
    t_a, v_a = ra.get_series("ram_in_use", mine.flat.server_ids[0])
    t_b, v_b = mine.get_series("ram_in_use", mine.flat.server_ids[0])
    assert v_a == v_b and np.allclose(t_a, t_b)
    assert mine.format_latency_stats() == ra.format_latency_stats()


@pytest.mark.parametrize("seed", [301, 305, 312, 327])
def test_sweep_row_equals_unmodified_reference_on_payload_for(seed):
    """A sweep point handed back to the reference: SweepSpec.payload_for(i) passes the reference's
    own SimulationPayload validation, and the UNMODIFIED reference actors run on it produce the
    clocks the engine produced for row i of the sweep."""
    import fuzz

    from asyncflow_b200.flatten import SweepSpec

    payload = fuzz.scenario(seed)
    flat = flatten(payload)
    n = 2
    spec = SweepSpec(flat, n, fuzz.sweep_columns(seed, payload, n))
    r = twin.run(flat, seed=SEED, replica_begin=0, n=n, sweep=spec, trace=n, clock_cap=100000, request_capacity=200000)
    for i in range(n):
        p = spec.payload_for(payload, i)
        ref = ref_harness.run_reference(p, seed=SEED, replica=i)          # validates with the reference schema
        k = int(r["stats"][i]["completed"])
        got = [tuple(x) for x in r["trace_clocks"][i, :k].tolist()]
        assert got == [tuple(w) for w in ref["clocks"]]
        assert dict(zip(flat.edge_ids, map(int, r["dropped"][i]))) == ref["edge_dropped"]


REFERENCE_YAMLS = ["examples/yaml_input/data/two_servers_lb.yml", "examples/yaml_input/data/event_inj_single_server.yml",
                   "examples/yaml_input/data/heavy_inj_single_server.yml", "examples/yaml_input/data/single_server.yml",
                   "examples/yaml_input/data/event_inj_lb.yml", "tests/integration/single_server/data/single_server.yml"]


@pytest.mark.parametrize("rel", REFERENCE_YAMLS)
def test_every_scenario_file_the_reference_ships_runs_identically(rel):
    """The reference's own example / test YAMLs, as they lie in /root/reference (read at test time, not
    copied): unmodified reference actors == engine state machine, clock for clock and series for series."""
    import yaml
    from pathlib import Path

    payload = yaml.safe_load((Path(ref_harness.REFERENCE_SRC).parent / rel).read_text())
    full = int(payload["sim_settings"].get("total_simulation_time", 3600))
    horizon = min(full, 40)
    payload["sim_settings"]["total_simulation_time"] = horizon
    for ev in payload.get("events") or []:                 # same timeline, compressed into the shortened horizon
        ev["start"]["t_start"] = float(ev["start"]["t_start"]) * horizon / full
        ev["end"]["t_end"] = float(ev["end"]["t_end"]) * horizon / full
    ref = ref_harness.run_reference(payload, seed=SEED, replica=2)
    mine = twin_results(payload, 2)
    assert [tuple(x) for x in mine.clocks.tolist()] == [tuple(c) for c in ref["clocks"]]
    assert mine.generated == ref["generated"] and mine.edge_dropped == ref["edge_dropped"]
    ma, mb = ref["analyzer"].get_sampled_metrics(), mine.get_sampled_metrics()
    for metric in ma:
        for ent in ma[metric]:
            assert list(ma[metric][ent]) == list(mb[metric][ent]), (metric, ent)
