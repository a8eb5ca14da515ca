"""Executed instructions of one kernel per source line (ncu source page joined with nvdisasm line info).

    cuobjdump -xelf all asyncflow_b200/_lib/libasyncflow_b200.so; nvdisasm -g -c af_engine.sm_100a.cubin > dis.txt
    ncu -i gpurun_out/prof_TAG.ncu-rep --page source --csv --print-source sass > sass.csv
    python tools/ncu_by_line.py dis.txt sass.csv [_Z14af_lane_kernelv] [--bucket 10] [--top 40]

The .so must be the one that was profiled (same instruction count, checked)."""
import argparse, collections, csv, re

ap = argparse.ArgumentParser()
ap.add_argument("dis"); ap.add_argument("sass"); ap.add_argument("kernel", nargs="?", default="_Z14af_lane_kernelv")
ap.add_argument("--bucket", type=int, default=10); ap.add_argument("--top", type=int, default=45)
ap.add_argument("--lines", default="", help="file:lo-hi -> dump the instructions of these source lines")
a = ap.parse_args()
lines = open(a.dis).read().split("\n")
start = end = None
for i, l in enumerate(lines):
    if re.match(r"\s*\.section\s+\.text\." + re.escape(a.kernel) + r"\b", l): start = i
    elif start is not None and end is None and re.match(r"\s*\.section\s+", l): end = i
cur, ins = None, []
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m2 = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m2: ins.append((int(m2.group(1), 16), m2.group(2).strip(), cur))
rows = list(csv.reader(open(a.sass)))
hdr, data = rows[1], rows[2:]
assert len(data) == len(ins), (len(data), len(ins))
ie, te, ss = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
tot_i = sum(int(r[ie]) for r in data); tot_t = sum(int(r[te]) for r in data); tot_s = sum(int(r[ss]) for r in data)
print(f"static {len(ins)}  warp-inst {tot_i:.4g}  thread-inst {tot_t:.4g}  lanes/inst {tot_t / tot_i:.2f}  samples {tot_s}")
b = collections.defaultdict(lambda: [0, 0, 0, 0])
for r, (addr, txt, cur) in zip(data, ins):
    f, ln = cur if cur else ("?", 0)
    k = (f, ln // a.bucket * a.bucket)
    b[k][0] += int(r[ie]); b[k][1] += int(r[te]); b[k][2] += int(r[ss]); b[k][3] += 1
    if a.lines:
        lf, rng = a.lines.split(":"); lo, hi = map(int, rng.split("-"))
        if f == lf and lo <= ln <= hi: print(f"  {ln:5d} {int(r[ie]):>12d} {int(r[te]) / max(int(r[ie]), 1):5.1f} {int(r[ss]):7d}  {txt}")
print(f"{'file:line':28s} {'inst%':>6s} {'lanes':>6s} {'samp%':>6s} {'static':>6s}")
for k in sorted(b, key=lambda k: -b[k][0])[: a.top]:
    v = b[k]
    print(f"{k[0] + ':' + str(k[1]):28s} {100 * v[0] / tot_i:6.1f} {v[1] / max(v[0], 1):6.1f} {100 * v[2] / tot_s:6.1f} {v[3]:6d}")
