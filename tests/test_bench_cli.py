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


def check_contract_line(d, n_gpus=1):
    """The keys the driver reads from our arm's JSON line (bench.py docstring)."""
    assert d["metric"] == "phonon drift-steps/sec" and d["unit"] == "drift-steps/s" and d["n_gpus"] == n_gpus
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["dtype"] == "f32" and d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] > 0
    assert d["config"]["workload"].startswith("synthetic 100-cell Si/Ge") and "l2" in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 1000 and r["achieved"] > 0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and "traffic" in r and r["algorithmic_bytes_per_drift_step"] == 64
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == "drift-steps/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != d["value"]
    c = d["clocks"]
    assert set(c) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    if n_gpus == 1 and "cpu_baseline" in d:
        b = d["cpu_baseline"]
        assert b["kind"] in ("reference", "port") and b["cores"] >= 1 and b["unit"] == "drift-steps/s" and b["sample"]


def test_committed_bench_lines_follow_the_contract():
    """profiles/r01_bench_n*.json are bench.py's own output lines (one B200 box each)."""
    for n in (1, 2, 4, 8):
        path = os.path.join(ROOT, "profiles", f"r01_bench_n{n}.json")
        check_contract_line(json.load(open(path)), n)
        assert ("cpu_baseline" in json.load(open(path))) == (n == 1)


@pytest.mark.gpu
def test_our_arm_prints_one_contract_line():
    r = _run(["--phonons", "2000000", "--steps", "1", "--warmup", "3", "--cpu-phonons-per-core", "2000"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    check_contract_line(d, 1)
    assert d["steps"] == 1 and d["warmup"] == 3 and d["config"]["phonons_per_gpu"] == 2_000_000
    assert d["cpu_baseline"]["value"] is None or d["cpu_baseline"]["value"] > 0
