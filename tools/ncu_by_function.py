"""Where the instructions go: aggregate an ncu source-page export per out-of-line function and per source line.

    ncu -i gpurun_out/prof.ncu-rep --page source --csv --print-source sass > /tmp/sass.csv
    python tools/ncu_by_function.py /tmp/sass.csv [library.so] [--lines 40] [--trim out.csv]

The kernel's helpers are __noinline__ device functions inside af_sim_kernel; ncu reports per SASS
instruction, nvdisasm names the sub-functions and (with -g) the source lines.  The library must be the
build the report was taken from (same SASS).
"""
from __future__ import annotations

import argparse
import bisect
import collections
import csv
import re
import subprocess
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def disassemble(lib: Path) -> str:
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", str(lib)], cwd=d, check=True, capture_output=True)
        cubin = next(Path(d).glob("*.cubin"))
        return subprocess.run(["nvdisasm", "-c", "-g", str(cubin)], check=True, capture_output=True, text=True).stdout


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("lib", nargs="?", default=str(ROOT / "asyncflow_b200" / "_lib" / "libasyncflow_b200.so"))
    ap.add_argument("--lines", type=int, default=40)
    ap.add_argument("--trim", default=None, help="write a trimmed copy of the export (address, sass, samples, executed, no_inst)")
    a = ap.parse_args()

    func_at: list[tuple[int, str]] = []
    line_of: dict[int, tuple[str, int]] = {}
    inside, cur_f, cur_l = False, None, None
    for ln in disassemble(Path(a.lib)).splitlines():
        if ln.startswith(".text._Z13af_sim_kernelv:"):
            inside, cur_f = True, "run_replica (+ inlined helpers)"
            continue
        if inside and ln.startswith(".text."):
            break
        if not inside:
            continue
        if ln.startswith("$"):
            name = ln.strip().rstrip(":").split("$")[-1]
            m = re.match(r"_ZN3af[cr]\d+([a-z_0-9]+?)E", name)
            cur_f = m.group(1) if m else name
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_l = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", ln)
        if m:
            off = int(m.group(1), 16)
            if cur_f:
                func_at.append((off, cur_f))
                cur_f = None
            line_of[off] = cur_l
    func_at.sort()
    starts = [o for o, _ in func_at]

    rows = list(csv.reader(open(a.csv)))
    hdr, data = rows[1], rows[2:]
    ia, isrc, ie, ismp, ini = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "# Samples", "stall_no_inst"))
    base = int(data[0][ia], 16)
    per_f: dict[str, list[int]] = {}
    per_l: collections.Counter = collections.Counter()
    tot_e = tot_s = 0
    main_end = starts[1] if len(starts) > 1 else 1 << 30
    for r in data:
        off = int(r[ia], 16) - base
        e, s, ni = int(r[ie] or 0), int(r[ismp] or 0), int(r[ini] or 0)
        f = func_at[bisect.bisect_right(starts, off) - 1][1]
        acc = per_f.setdefault(f, [0, 0, 0, 0])
        acc[0] += e; acc[1] += s; acc[2] += ni; acc[3] += 1
        tot_e += e; tot_s += s
        if off < main_end:
            per_l[line_of.get(off)] += e
    print(f"warp instructions executed {tot_e:.4e}, stall samples {tot_s}")
    print(f"{'function':34s} {'SASS':>5s} {'executed':>9s} {'samples':>8s} {'no_inst share of its samples':>30s}")
    for f, (e, s, ni, n) in sorted(per_f.items(), key=lambda kv: -kv[1][0]):
        if e / tot_e < 0.0005:
            continue
        print(f"{f:34s} {n:5d} {e / tot_e * 100:8.2f}% {s / tot_s * 100:7.2f}% {ni / max(s, 1) * 100:29.1f}%")
    src = {p.name: p.read_text().split("\n") for p in (ROOT / "asyncflow_b200" / "csrc").glob("*")}
    print(f"\nrun_replica by source line (top {a.lines}; share of ALL executed instructions)")
    for key, e in per_l.most_common(a.lines):
        f, l = key if key else ("?", 0)
        text = src[f][l - 1].strip()[:100] if f in src and 0 < l <= len(src[f]) else ""
        print(f"{e / tot_e * 100:6.2f}%  {f}:{l}  {text}")
    if a.trim:
        with open(a.trim, "w", newline="") as fh:
            w = csv.writer(fh)
            w.writerow(["offset", "sass", "samples", "executed", "stall_no_inst"])
            for r in data:
                w.writerow([hex(int(r[ia], 16) - base), r[isrc], r[ismp], r[ie], r[ini]])


if __name__ == "__main__":
    main()
