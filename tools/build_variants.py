"""Build the experimental variants of the engine next to the product library.

    python tools/build_variants.py            # all variants
    ASYNCFLOW_B200_LIB=asyncflow_b200/_lib/libasyncflow_b200_memo.so python tools/quick_bench.py ...

Each variant is the same source with extra -D flags (af_core.cuh documents them); results are
bit-identical by construction and tests/test_variants.py checks that on the CPU twin.  They exist so a
GPU session can A/B them against the product build in ONE gpurun call (tools/ab_variants.sh).
"""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as G  # noqa: E402

VARIANTS = {
    "predraw": ["-DAF_PREDRAW"],                 # lane-parallel memoised edge variates
    "pregen": ["-DAF_PREGEN"],                   # lane-parallel memoised inter-arrival logs
    "memo": ["-DAF_PREDRAW", "-DAF_PREGEN"],     # both
    "sorted": ["-DAF_SORTED_POOL"],              # sorted 32-entry front ring of the pending-event pool
    "pin": ["-DAF_PIN_ACTIVE"],                  # requests being served stay in the shared-memory tier
    "all": ["-DAF_PREDRAW", "-DAF_PREGEN", "-DAF_SORTED_POOL"],
    "all4": ["-DAF_PREDRAW", "-DAF_PREGEN", "-DAF_SORTED_POOL", "-DAF_PIN_ACTIVE"],
    "all_mb6": ["-DAF_PREDRAW", "-DAF_PREGEN", "-DAF_SORTED_POOL", "-DAF_MIN_BLOCKS=6"],   # 80 registers, 24 warps/SM
}


def build(names=None, verbose: bool = False) -> list[Path]:
    out = []
    for name in names or VARIANTS:
        so = G.LIB.with_name(f"libasyncflow_b200_{name}.so")
        cmd = [G._nvcc(), *G.NVCC_FLAGS, *VARIANTS[name], *(["-Xptxas", "-v"] if verbose else []),
               "-o", str(so), str(G.CSRC / "af_engine.cu")]
        r = subprocess.run(cmd, check=True, cwd=ROOT, capture_output=True, text=True)
        if verbose:
            lines = r.stderr.splitlines()
            for i, ln in enumerate(lines):
                if "Function properties for _Z13af_sim_kernelv" in ln:
                    print(f"[{name}] " + " | ".join(x.strip() for x in lines[i + 1:i + 3]))
        out.append(so)
    return out


if __name__ == "__main__":
    for p in build(sys.argv[1:] or None, verbose=True):
        print(p)
