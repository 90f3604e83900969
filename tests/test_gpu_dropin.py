"""The drop-in, compiled and run: oracle/_ref/psim_ref_b200 is the reference's OWN program - its main.cpp,
JSON loader, mesh validation, material tables, run epilogue and exporter, all unmodified - with psim/src/modelSimulator.cpp
replaced by oracle/ref_shim/modelSimulatorB200.cpp, which implements ModelSimulator's public methods
(modelSimulator.h:12-41) over the C ABI of include/psim_b200.h.  The ss_*.txt tables that program writes must agree with
the fixtures made from 16 seeds of the unmodified reference."""
import os
import subprocess
import sys

import numpy as np
import pytest

from psim_b200 import configs
from tests import common as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "psim_ref_b200")


def read_ss(path):
    lines = open(path).read().splitlines()
    return lines[0], np.array([[float(x) for x in ln.split()] for ln in lines[1:] if ln.strip()])


def test_dropin_binary_links_the_c_abi_library():
    """(CPU) the binary exists wherever the reference tree was available to build it, and its only non-system dependency is
    libpsim_b200.so."""
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/psim_ref_b200 not built (no reference tree on this machine)")
    out = subprocess.run(["ldd", BIN], capture_output=True, text=True).stdout
    assert "libpsim_b200.so" in out


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["linear_demo", "sides_ss", "sige"])
def test_reference_program_with_the_gpu_underneath_matches_the_reference(name, tmp_path):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/psim_ref_b200 not built")
    gold = T.golden(name)
    tables = []
    for seed in range(1, 9):
        path = configs.save(T.case_model(name), str(tmp_path / f"{name}.json"))
        r = subprocess.run([BIN, path], capture_output=True, text=True, timeout=600, env=dict(os.environ, PSIM_SEED=str(seed)))
        assert r.returncode == 0 and "Run: 1" in r.stdout and "psim_b200:" not in r.stderr, (r.stdout, r.stderr)
        header, table = read_ss(str(tmp_path / f"ss_{name}.txt"))
        assert header.startswith("Steady State Results from ") and " over 1 runs" in header
        tables.append(table)
    runs = [{"out6": t} for t in tables]
    assert tables[0].shape == gold["out6_mean"].shape
    z = T.welch_z(runs, gold, "out6")
    T.assert_parity(z[:, 0], f"{name}: temperature column of the reference program's ss table, GPU underneath")
    T.assert_parity(z[:, 2], f"{name}: x-flux column")
    T.assert_parity(z[:, 4], f"{name}: y-flux column")


@pytest.mark.gpu
def test_dropin_reports_errors_the_reference_way(tmp_path):
    """An exception of the GPU path surfaces through the reference's own main (print e.what(), continue): a device index
    that does not exist."""
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/psim_ref_b200 not built")
    path = configs.save(configs.linear(num_phonons=10_000).to_dict(), str(tmp_path / "m.json"))
    r = subprocess.run([BIN, path], capture_output=True, text=True, timeout=120, env=dict(os.environ, PSIM_DEVICE="99"))
    assert r.returncode == 0 and "psim_b200: device index out of range" in r.stderr and "done" in r.stdout
    assert not os.path.exists(str(tmp_path / "ss_m.txt"))
