"""Model-file generators (psim_b200/configs.py) against the reference's shipped files (when the tree is mounted)."""
import json
import os

import pytest

from psim_b200 import configs

REF = "/root/reference/psim_python/json"


def _same(a, b, path=""):
    if isinstance(a, dict):
        assert isinstance(b, dict) and set(a) == set(b), path
        for k in a:
            _same(a[k], b[k], path + "/" + k)
    elif isinstance(a, list):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{path}[{i}]")
    else:
        assert a == b, (path, a, b)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_generators_reproduce_shipped_models():
    _same(configs.linear().to_dict(), json.load(open(f"{REF}/linear_demo.json")))
    _same(configs.linear_sides().to_dict(), json.load(open(f"{REF}/linear_sides_demo_ss.json")))
    _same(configs.linear_sides(sim_type=1, step_interval=4).to_dict(), json.load(open(f"{REF}/linear_sides_demo_per.json")))
    _same(configs.linear_sides(sim_type=2, step_interval=4, start_time=0.1, duration=0.15).to_dict(),
          json.load(open(f"{REF}/linear_sides_demo_trans.json")))


def test_si_ge_grid_shape():
    m = configs.si_ge_grid(num_phonons=1000).to_dict()
    assert len(m["cells"]) == 100 and len(m["sensors"]) == 50 and len(m["emit_surfaces"]) == 10
    mats = [s["material"] for s in m["sensors"]]
    assert mats[:25] == ["Silicon"] * 25 and mats[25:] == ["Germanium"] * 25  # silicon cells first
    assert m["settings"]["t_eq"] == 300.0 and m["settings"]["sim_type"] == 0
