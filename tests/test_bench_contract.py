"""CPU test of the benchmark contract: `bench.py --impl reference` (the reference's own CPU path through
oracle/_ref, or the oracle port) prints exactly ONE JSON line on stdout carrying the keys the driver reads, both
as a plain process and under torchrun with two ranks (rank 0 alone prints; the other exits 0 without work)."""
from __future__ import annotations

import json
import os
import subprocess
import sys

from conftest import ROOT

REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _check(stdout: str, n_gpus: int) -> None:
    lines = [ln for ln in stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["impl"] == "reference" and d["n_gpus"] == n_gpus and d["gpu_launches"] == 0
    assert d["unit"] == "ciphertexts/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert "workload" in d["config"] and "n=4096" in d["config"]["workload"]


def test_reference_arm_single_process():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-800:]
    _check(r.stdout, 1)


def test_reference_arm_under_torchrun():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577", os.path.join(ROOT, "bench.py"), "--impl",
                        "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-800:]
    _check(r.stdout, 2)
