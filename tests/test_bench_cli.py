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


def check_contract_line_r01(d, n_gpus=1):
    """Round 1's line (profiles/r01_bench_n*.json): weak scaling, the 64 B x drift-steps accounting as `roofline`."""
    assert d["metric"] == "phonon drift-steps/sec" and d["unit"] == "drift-steps/s" and d["n_gpus"] == n_gpus
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["dtype"] == "f32" and d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 1000 and r["achieved"] > 0


def check_contract_line(d, n_gpus=1, models=None, kinked=True):
    """The keys the driver reads from our arm's JSON line (bench.py docstring)."""
    assert d["metric"] == "phonon drift-steps/sec" and d["unit"] == "drift-steps/s" and d["n_gpus"] == n_gpus
    assert d["higher_is_better"] is True and d["scaling"] == "strong" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["dtype"] == "f32" and d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] > 0
    assert d["config"]["workload"].startswith("synthetic 100-cell Si/Ge") and "l2" in d["config"]
    assert d["config"]["phonons_per_gpu"] * n_gpus <= d["config"]["phonons_total"] < (d["config"]["phonons_per_gpu"] + 1) * n_gpus
    r = d["roofline"]
    assert r["bound"] == "issue" and r["unit"] == "warp-inst/s" and r["peak"] > 1e11 and "traffic" in r and r["segments_per_s"] > 0
    h = r["hbm"]
    assert h["bound"] == "hbm" and h["unit"] == "GB/s" and h["peak"] > 1000 and h["algorithmic_bytes_per_drift_step"] == 64
    if r["frac"] is None:  # no ncu capture of exactly this source tree and configuration: nothing stale is reported
        assert r["achieved"] is None and r["traffic"] is None and h["achieved"] is None and h["frac"] is None
    else:
        assert 0 < r["frac"] <= 1.0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert 0 < h["frac"] <= 1.0 and abs(h["frac"] - h["achieved"] / h["peak"]) < 1e-9 and r["traffic"] > 0
        assert r["warp_instructions_per_segment"] > 1
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == "drift-steps/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != d["value"]
    c = d["clocks"]
    assert set(c) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    names = ["si_ge_grid"] + (["kinked"] if kinked else [])
    for name in names:
        k = d["strong"][name]
        for key in ("tallies", "energy", "flux", "emitted_counts", "cell_histogram_mid_run"):
            assert len(k[key]) == 16 and int(k[key], 16) >= 0
        assert k["phonons_emitted"] > 0 and k["drift_steps"] > 0 and k["allreduce_bytes_per_job"] > 0
    assert (d["weak"] is None) == (n_gpus == 1)
    if models if models is not None else n_gpus == 1:
        ms = {m["model"]: m for m in d["models"]}
        assert {"linear_demo", "linear_sides_demo_ss", "linear_sides_demo_per", "linear_sides_demo_trans", "si_ge_grid_1e8"} <= set(ms)
        for m in ms.values():
            assert m["ms_e2e"] > 0 and m["kernel_ms"] > 0 and m["drift_steps_per_s"] > 0 and m["segments_per_s"] > 0
        assert ms["linear_demo"]["reference_header_s"] == 18.1
    if n_gpus == 1 and "cpu_baseline" in d:
        b = d["cpu_baseline"]
        assert b["kind"] in ("reference", "port") and b["cores"] >= 1 and b["unit"] == "drift-steps/s" and b["sample"]


def test_committed_bench_lines_follow_the_contract():
    """profiles/r0*_bench_n*.json are bench.py's own output lines (one B200 box each); the strong-scaling lines of one
    round must carry the same checksums for every number of GPUs."""
    for n in (1, 2, 4, 8):
        check_contract_line_r01(json.load(open(os.path.join(ROOT, "profiles", f"r01_bench_n{n}.json"))), n)
    sums = {}
    for n in (1, 2, 4, 8):
        path = os.path.join(ROOT, "profiles", f"r02_bench_n{n}.json")
        if not os.path.exists(path):
            continue
        d = json.load(open(path))
        check_contract_line(d, n)
        assert ("cpu_baseline" in d) == (n == 1)
        for name, k in d["strong"].items():
            key = tuple(k[x] for x in ("tallies", "emitted_counts", "cell_histogram_mid_run", "drift_steps"))
            assert sums.setdefault(name, key) == key, (name, n)


@pytest.mark.gpu
def test_our_arm_prints_one_contract_line():
    r = _run(["--phonons", "2000000", "--steps", "1", "--warmup", "3", "--cpu-phonons-per-core", "2000", "--no-kinked", "--no-models"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    check_contract_line(d, 1, models=False, kinked=False)
    assert d["steps"] == 1 and d["warmup"] == 3 and d["config"]["phonons_per_gpu"] == 2_000_000
    assert d["cpu_baseline"]["value"] is None or d["cpu_baseline"]["value"] > 0
