"""The GPU-session tools and examples, dry-run on the twin-backed engine so they do not rot between
GPU sessions (Python-level only: argument plumbing, the SweepRunner calls, the printed summary)."""

from __future__ import annotations

import runpy
import sys

import pytest
from helpers import ROOT
from twin_engine import TwinEngine


@pytest.fixture()
def twin_backed(monkeypatch):
    import asyncflow_b200
    import asyncflow_b200.runner as R
    monkeypatch.setattr(R, "Engine", TwinEngine)
    monkeypatch.setattr(asyncflow_b200, "Engine", TwinEngine)


@pytest.mark.parametrize("args", [
    "--scenario c3_lb_two_servers.yml --replicas 12 --horizon 5 --reps 1",
    "--scenario c1_my_service.yml --replicas 12 --horizon 5 --reps 1 --sweep users --balance",
    "--scenario c4_lb8_events.yml --replicas 4 --horizon 10 --reps 1 --sweep none --no-metrics",
])
def test_quick_bench(twin_backed, monkeypatch, capsys, args):
    monkeypatch.setattr(sys, "argv", ["quick_bench.py", *args.split()])
    runpy.run_path(str(ROOT / "tools" / "quick_bench.py"), run_name="__main__")
    out = capsys.readouterr().out
    assert "compl/s" in out and "overflow 0.0" in out


def test_ab_lane_lib(twin_backed, monkeypatch, capsys):
    import json
    monkeypatch.setattr(sys, "argv", ["ab_lane_lib.py", "--replicas", "6", "--horizon", "4", "--reps", "1"])
    runpy.run_path(str(ROOT / "tools" / "ab_lane_lib.py"), run_name="__main__")
    row = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert row["checksum"][0] > 0 and row["flags"] == 0 and row["ms"] > 0


def test_drilldown_example(twin_backed, monkeypatch, capsys):
    monkeypatch.setattr(sys, "argv", ["sweep_users_drilldown.py", "24"])
    runpy.run_path(str(ROOT / "examples" / "sweep_users_drilldown.py"), run_name="__main__")
    out = capsys.readouterr().out
    assert "replayed" in out and "as a reference payload" in out
