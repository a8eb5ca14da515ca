"""profiles/kernel_metrics.json from the metrics pass of tools/profile_gpu.sh.

    python tools/ncu_issue_summary.py r02a [--config c3]

Reads gpurun_out/metrics_<tag>.csv (ncu --csv, one launch of the dominant kernel) and the bench line the same command
printed (gpurun_out/b_metrics_<tag>.log: events in that launch), copies both next to the other evidence under profiles/
and records, per "<config>:<kernel>": DRAM bytes per launch, executed warp instructions per timed event, issue-active %,
the stall ratios.  bench.py reads the JSON for roofline.traffic / roofline.issue; every figure in it can be recomputed
from the two copied files.
"""
from __future__ import annotations

import argparse
import csv
import json
import shutil
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--config", default="c3")
    a = ap.parse_args()
    src = ROOT / "gpurun_out" / f"metrics_{a.tag}.csv"
    log = ROOT / "gpurun_out" / f"b_metrics_{a.tag}.log"
    rows = [r for r in csv.reader(src.read_text().splitlines()) if len(r) > 5]
    head = next(i for i, r in enumerate(rows) if "Metric Name" in r)
    H = rows[head]
    name_i, val_i, kern_i = H.index("Metric Name"), H.index("Metric Value"), H.index("Kernel Name")
    m, kernel = {}, None
    for r in rows[head + 1:]:
        kernel = r[kern_i].split("(")[0].split("<")[0]
        try:
            m[r[name_i]] = float(r[val_i].replace(",", ""))
        except ValueError:
            continue                                  # "n/a": metric not available on this chip
    line = next(json.loads(ln) for ln in log.read_text().splitlines() if ln.startswith("{"))
    steps = line["steps"]
    events = line["events_per_s"] * line["ms_per_step"] / 1e3 / line["n_gpus"]       # timed events of one launch
    out_p = ROOT / "profiles" / "kernel_metrics.json"
    out = json.loads(out_p.read_text()) if out_p.exists() else {}
    dst_csv = ROOT / "profiles" / f"{a.tag}_ncu_metrics_{kernel}.csv"
    dst_log = ROOT / "profiles" / f"{a.tag}_ncu_metrics_bench_line.json"
    shutil.copy(src, dst_csv)
    dst_log.write_text(json.dumps(line) + "\n")
    stalls = {k.split("issue_stalled_")[1].split("_per_issue")[0]: v for k, v in m.items() if "issue_stalled_" in k}
    out[f"{a.config}:{kernel}"] = {
        "dram_bytes_per_launch": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"],
        "warp_inst_per_launch": m["smsp__inst_executed.sum"],
        "events_per_launch": events,
        "warp_inst_per_event": m["smsp__inst_executed.sum"] / events,
        "thread_inst_per_warp_inst": m["smsp__thread_inst_executed.sum"] / m["smsp__inst_executed.sum"],
        "issue_active_pct": m["smsp__issue_active.avg.pct_of_peak_sustained_active"],
        "warps_active_pct": m.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": m.get("launch__registers_per_thread"),
        "kernel_ms_under_ncu": m["gpu__time_duration.sum"] / 1e6,
        "stalled_warps_per_issue": stalls,
        "source": f"profiles/{dst_csv.name} + profiles/{dst_log.name} (tools/profile_gpu.sh pass 2, {steps} step)",
    }
    out_p.write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out[f"{a.config}:{kernel}"], indent=1))


if __name__ == "__main__":
    main()
