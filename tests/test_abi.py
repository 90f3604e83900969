"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares; nothing here makes a
compute call.  Without a GPU the library must refuse loudly (no CPU path)."""
import ctypes as C
import os
import re

import pytest

from psim_b200 import lib as psim
from tests import common as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for header in ("psim_b200.h", "psim_host.h"):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b(psim_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_every_declared_symbol_is_exported_and_bound(built_library):
    lib = C.CDLL(built_library)
    names = declared_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ but not exported"
    assert set(names) == set(psim.SYMBOLS), set(names) ^ set(psim.SYMBOLS)


def test_struct_layouts_match_headers(built_library):
    # sizes the C compiler gives the public structs (include/psim_b200.h) vs the ctypes mirrors
    assert C.sizeof(psim.Material) == 64
    assert C.sizeof(psim.Sensor) == 24
    assert C.sizeof(psim.SubSurface) == 48
    assert C.sizeof(psim.Cell) == 88
    assert C.sizeof(psim.Emitter) == 48
    assert C.sizeof(psim.Source) == 24
    assert C.sizeof(psim.Table) == 16
    assert C.sizeof(psim.ModelDesc) == 112  # + step_sensors (round 2)


def test_no_device_is_a_loud_error(built_library):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = T.load_model(T.case_model("linear_demo"), num_phonons=1000)
    m.prepare()
    with pytest.raises(psim.PsimError) as e:
        psim.GpuSimulator(m.describe(), 0)
    assert e.value.code == -2 and "no CPU path" in str(e.value)
    with pytest.raises(psim.PsimError) as e:
        m.run(device=0, seed=1)
    assert e.value.code == -2


def test_product_does_not_reach_into_the_oracle():
    """oracle/ and tests/emu are checkers: the package sources must not reference them."""
    pkg = os.path.join(ROOT, "psim_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                if f == "build.py":
                    continue  # the build recipe names the test-only emulation target; it loads nothing
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle" not in text.replace("Oracle", ""), os.path.join(base, f)
                assert "libpsim_emu" not in text, os.path.join(base, f)
