"""Device parity of the engine's build variants (tools/build_variants.py) -- opt-in.

Skipped unless AF_TEST_VARIANTS=1 and the variant libraries exist: the product build is what the GPU
suite checks; a variant is checked when a session wants to promote it (tools/ab_variants.sh does the same
before timing).  Each variant runs in its own process because the library is bound at first import."""

from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
LIBDIR = ROOT / "asyncflow_b200" / "_lib"

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("AF_TEST_VARIANTS") != "1", reason="set AF_TEST_VARIANTS=1 to check build variants")]


@pytest.mark.parametrize("variant", ["predraw", "pregen", "memo", "sorted", "all"])
def test_variant_is_bit_exact_on_the_device(variant):
    lib = LIBDIR / f"libasyncflow_b200_{variant}.so"
    if not lib.exists():
        pytest.skip(f"{lib.name} not built (python tools/build_variants.py)")
    env = dict(os.environ, ASYNCFLOW_B200_LIB=str(lib))
    p = subprocess.run([sys.executable, str(ROOT / "tools" / "check_variant_gpu.py")], env=env, cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "OK:" in p.stdout
