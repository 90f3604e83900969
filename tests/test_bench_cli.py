"""bench.py as the driver calls it, as far as that goes without a GPU: the reference arm (the reference's own CPU code
from oracle/_ref, or the oracle port when that binary is absent) prints ONE JSON line with the contract's keys; our arm
refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                          env=dict(os.environ, **(env or {})), cwd=ROOT)


def test_reference_arm_prints_one_contract_line(built_library):
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-phonons-per-core", "3000"])
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "phonon drift-steps/sec" and d["unit"] == "drift-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("synthetic 100-cell Si/Ge")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "drift-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "psim_ref")):
        assert cb["kind"] == "reference"


def test_reference_arm_other_ranks_print_nothing(built_library):
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-phonons-per-core", "1000"],
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_needs_a_gpu(built_library):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = _run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr
