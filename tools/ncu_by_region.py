"""Executed instructions of af_lane_kernel per code region and per timed event.

    python tools/ncu_by_region.py dis.txt sass.csv gpurun_out/b_full_TAG.log

Regions are found from markers in asyncflow_b200/csrc/af_lane.cuh, so they follow the source as it moves."""
import collections, csv, json, re, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
dis, sass, log = sys.argv[1:4]
src = (ROOT / "asyncflow_b200/csrc/af_lane.cuh").read_text().split("\n")
def find(pat):
    for i, l in enumerate(src, 1):
        if re.search(pat, l): return i
    raise SystemExit("marker not found: " + pat)
marks = [("accessors smem", r"^struct Mem"), ("accessors global/tier", r"^AFL_IN unsigned char\* g128p"), ("request/event wrappers", r"^AFL_IN void rq_load"),
         ("heap", r"^AFL_IN void ev_get"), ("now-queue", r"^AFL_IN uint64_t nq_ld"), ("gen_next_gap", r"^AFL_IN bool gen_next_gap"),
         ("gauges/ticks", r"^AFL_IN void gauge_touch"), ("stores/walks/injection", r"^AFL_IN bool q_empty"), ("complete", r"^AFL_IN void complete"),
         ("start/write-back", r"^AFL_IN void start_replica"), ("loop: lifecycle", r"phase: lifecycle"), ("loop: gap", r"phase: the generator's next timeout"),
         ("loop: pick", r"phase: pick the next"), ("loop: ticks", r"phase: collector ticks"), ("loop: decode", r"phase: decode"),
         ("loop: node", r"phase: a node's consumer"), ("loop: steps", r"phase: the `for step"), ("loop: send", r"phase: EdgeRuntime.transport"),
         ("loop: timer", r"phase: schedule the timeout")]
bounds = [(n, find(p)) for n, p in marks] + [("end", 10**9)]
line = json.loads([l for l in open(log) if l.startswith("{")][0])
ev = line["events_per_s"] * line["ms_per_step"] / 1e3
lines = open(dis).read().split("\n")
start = end = None
for i, l in enumerate(lines):
    if re.match(r"\s*\.section\s+\.text\._Z14af_lane_kernelv\b", l): start = i
    elif start is not None and end is None and re.match(r"\s*\.section\s+", l): end = i
cur, ins = None, []
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+", l): ins.append(cur)
rows = list(csv.reader(open(sass))); hdr, data = rows[1], rows[2:]
assert len(data) == len(ins), (len(data), len(ins))
ie, te = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
agg = collections.defaultdict(lambda: [0, 0, 0])
for r, cur in zip(data, ins):
    f, ln = cur if cur else ("?", 0)
    g = "other"
    if f == "af_rng.cuh": g = "rng (af_rng.cuh)"
    elif f == "af_lane.cuh":
        for (n, lo), (_, hi) in zip(bounds, bounds[1:]):
            if lo <= ln < hi: g = n; break
    a = agg[g]; a[0] += int(r[ie]); a[1] += int(r[te]); a[2] += 1
print(f"timed events in the launch: {ev:.4g}")
print(f"{'region':26s} {'warp-inst/ev':>12s} {'thread-inst/ev':>14s} {'lanes':>6s} {'static':>6s}")
for g in [n for n, _ in marks] + ["rng (af_rng.cuh)", "other"]:
    a = agg[g]
    print(f"{g:26s} {a[0] / ev:12.2f} {a[1] / ev:14.1f} {a[1] / max(a[0], 1):6.1f} {a[2]:6d}")
ti = sum(a[0] for a in agg.values()); tt = sum(a[1] for a in agg.values())
print(f"{'total':26s} {ti / ev:12.2f} {tt / ev:14.1f} {tt / ti:6.1f} {len(ins):6d}")
