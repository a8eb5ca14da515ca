"""Shared fixtures.  `gpu` marks tests that need a real B200 (driver runs them with -m gpu)."""

from __future__ import annotations

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "oracle", ROOT / "oracle" / "simpy_shim", ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

SEED = 0xA5F10


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


@pytest.fixture(scope="session", autouse=True)
def _native_code_is_built():
    """A fresh checkout has no .so files (they are git-ignored): build them once per session."""
    lib = ROOT / "asyncflow_b200" / "_lib" / "libasyncflow_b200.so"
    if not lib.exists():
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def scenarios_dir() -> Path:
    return ROOT / "tests" / "scenarios"
