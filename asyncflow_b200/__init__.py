"""asyncflow_b200 -- B200-native batched discrete-event engine for AsyncFlow scenarios.

One hot path of AsyncFlow (SimPy ``Environment.step()`` driving the actors,
reference ``src/asyncflow/runtime/simulation_runner.py:349-376``) re-built as
hand-written sm_100a CUDA behind the reference's own surface:

* IN : a ``SimulationPayload`` / YAML dict       -> :func:`flatten`
* RUN: :class:`GpuSimulationRunner` (one replica, mirrors ``SimulationRunner``)
       :class:`SweepRunner` (10^4-10^6 replicas of a parameter sweep)
* OUT: ``ResultsAnalyzer``-compatible results     -> :mod:`asyncflow_b200.results`

There is no CPU fallback: without the CUDA library / a B200 the runners raise
:class:`EngineUnavailable`.
"""

from ._capi import EngineUnavailable
from .engine import Engine, EngineError
from .flatten import FlatScenario, SweepSpec, balanced_order, flatten
from .results import ReplicaResults, SweepResults, series_bands
from .runner import GpuSimulationRunner, SweepRunner

__all__ = [
    "Engine", "EngineError", "EngineUnavailable", "FlatScenario", "SweepSpec", "balanced_order", "flatten",
    "GpuSimulationRunner", "SweepRunner", "ReplicaResults", "SweepResults", "series_bands",
]
