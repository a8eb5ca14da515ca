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


def test_missing_library_is_an_error_not_a_fallback():
    """No CPU path behind the product API: without the CUDA library the first use raises."""
    import subprocess
    import sys
    code = ("import os; os.environ['ASYNCFLOW_B200_LIB'] = '/nonexistent/libasyncflow_b200.so'\n"
            "from asyncflow_b200 import Engine, EngineUnavailable\n"
            "try:\n    Engine(0)\nexcept EngineUnavailable as e:\n    print('raised', 'no CPU fallback' in str(e))\n")
    p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert p.stdout.strip() == "raised True", p.stdout + p.stderr


def test_product_package_never_reaches_into_the_oracle_or_the_tests():
    """oracle/ and tests/ are checkers: nothing under asyncflow_b200/ may import or locate them."""
    import re
    bad = re.compile(r"^\s*(from|import)\s+(des_port|afrng|afrng_c|ref_harness|simpy|twin|fuzz|helpers)\b|sys\.path|oracle[/\\]|host_twin")
    for path in sorted((ROOT / "asyncflow_b200").rglob("*.py")):
        for i, line in enumerate(path.read_text().splitlines(), 1):
            code = line.split("#", 1)[0]
            assert not bad.search(code) or "``" in line or '"""' in line or "oracle/afrng.py" in line, f"{path}:{i}: {line.strip()}"
    for path in sorted((ROOT / "asyncflow_b200" / "csrc").glob("*")):
        text = path.read_text()
        assert "#include \"../../oracle" not in text and "#include \"../../tests" not in text, path
