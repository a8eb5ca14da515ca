"""The C-ABI shared library: it loads, exports every symbol the header declares,
and refuses to run without a GPU (no CPU fallback)."""

from __future__ import annotations

import ctypes as C
import re

import pytest
from helpers import ROOT

from asyncflow_b200 import _capi as K

HEADER = ROOT / "include" / "asyncflow_b200.h"


def declared_functions() -> list[str]:
    src = HEADER.read_text()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(af_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = K.load()
    names = declared_functions()
    assert set(names) == set(K.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.af_abi_version() == K.AF_ABI_VERSION


def test_struct_sizes_match_the_header_layout():
    import twin
    structs = [K.AfEdge, K.AfServer, K.AfEndpoint, K.AfStep, K.AfSpikeMark, K.AfOutageMark, K.AfScenario,
               K.AfSweepColumn, K.AfSweep, K.AfOptions, K.AfReplicaStats]
    for i, st in enumerate(structs):
        assert C.sizeof(st) == twin.lib().af_twin_sizeof(i), st.__name__
    assert C.sizeof(K.AfReplicaStats) == 88 == K.STATS_DTYPE.itemsize


def test_engine_create_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import asyncflow_b200 as af
    with pytest.raises(af.EngineUnavailable, match="no CPU fallback"):
        af.Engine(0)
    with pytest.raises(af.EngineUnavailable):
        af.GpuSimulationRunner(simulation_input={"rqs_input": {"id": "g", "avg_active_users": {"mean": 1},
                               "avg_request_per_minute_per_user": {"mean": 1}},
                               "topology_graph": {"nodes": {"client": {"id": "c"}, "servers": [
                                   {"id": "s", "server_resources": {}, "endpoints": [{"endpoint_name": "e", "steps": [
                                       {"kind": "io_wait", "step_operation": {"io_waiting_time": 0.1}}]}]}]},
                                   "edges": [{"id": "a", "source": "g", "target": "c", "latency": {"mean": 0.1}},
                                             {"id": "b", "source": "c", "target": "s", "latency": {"mean": 0.1}},
                                             {"id": "d", "source": "s", "target": "c", "latency": {"mean": 0.1}}]},
                               "sim_settings": {"total_simulation_time": 5}}).run()
