"""`bench.py --impl reference`: the reference's own CPU implementation of the path (`bioframe.ops._overlap_intidxs`
from baseline/_ref, or the numpy port when that is absent) on every host core -- the arm the driver launches next to
ours.  Needs no GPU, so its contract is checked here at a small size: one JSON line, our arm's metric / unit /
direction, `impl: reference`, a `cpu_baseline` describing the run, an `e2e` that repeats the value with zero copies;
under torchrun rank 0 alone works and prints."""
import json
import os
import subprocess
import sys

from conftest import ROOT

SMALL = ["--impl", "reference", "--n1", "60000", "--n2", "6000", "--steps", "1", "--warmup", "1"]


def _lines(cmd):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return [json.loads(ln) for ln in out.stdout.splitlines() if ln.startswith("{")]


def _check(line, n_gpus):
    assert line["impl"] == "reference" and line["metric"] == "bf.overlap input intervals/s"
    assert line["unit"] == "intervals/s" and line["higher_is_better"] is True and line["n_gpus"] == n_gpus
    assert line["steps"] == 1 and line["warmup"] == 1 and line["value"] > 0 and line["ms_per_step"] > 0
    base = line["cpu_baseline"]
    assert base["kind"] in ("reference", "port") and base["cores"] >= 1 and base["value"] == line["value"]
    assert "chromosomes" in base["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert line["config"]["rows_out"] >= line["config"]["n1_per_gpu"] * 0.99  # a left join keeps every query


def test_reference_arm_prints_one_contract_line():
    lines = _lines([sys.executable, "bench.py", *SMALL])
    assert len(lines) == 1
    _check(lines[0], 1)


def test_reference_arm_under_torchrun_runs_on_rank_0_only():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29653", "bench.py", "--gpus", "2", *SMALL]
    lines = _lines(cmd)
    assert len(lines) == 1
    _check(lines[0], 2)
