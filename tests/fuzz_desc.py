"""Damaged model DESCRIPTIONS and source lists through the validation the C ABI applies before anything reaches the
device (psim_b200/csrc/flatten.cpp: flatten_model, plan_births), followed by a run of 3000 phonons through the CPU-side
emulation of the device functions whenever the description is accepted: indices out of range, wrapped 32-bit ranges,
non-finite coordinates and times, absurd counts.  The answer must be an error code or a finished run - never a crash or a
hang.  Run by tests/test_emu.py in a subprocess: usage  python tests/fuzz_desc.py <seed> <iterations>"""
import ctypes as C, random, sys, json
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from psim_b200 import configs, lib as psim
from tests import common as T
rng = random.Random(int(sys.argv[1]))
# FUZZ_STEADY=1: a steady-state model instead, whose unrecorded passes fly the lattice image (device_types.h)
if os.environ.get("FUZZ_STEADY"):
    model = T.load_model(configs.linear_sides(num_cells=20, y=20, num_phonons=3000).to_dict())
else:
    model = T.load_model(configs.linear_sides(num_cells=20, y=20, num_phonons=3000, sim_type=2, step_interval=4, start_time=0.1, duration=0.15).to_dict())
model.prepare()
lib = T.emu_lib()
desc = model.describe().contents
info = model.info
S, R, M = info.num_sensors, info.recorded_steps, info.measurement_steps
_src, _n = model.sources(1)
def run(nsrc=None, src=None):
    if nsrc is None: nsrc, src = _n, _src
    # (the caller sizes the tally arrays by the description it passes: R = measurement_steps - step_adjustment, which the
    # mutations below may raise up to M)
    e = np.zeros((S, M), dtype=np.int32); f = np.zeros((S, M, 2)); fx = np.zeros((S, M, 2), dtype=np.int64)
    steps, events = C.c_uint64(), C.c_uint64(); err = C.create_string_buffer(512)
    return lib.psim_emu_run(C.byref(desc), src, nsrc, 1, 0, 1, 16, e.ctypes.data, f.ctypes.data, fx.ctypes.data, C.byref(steps), C.byref(events), None, None, err, 512), err.value
print("baseline", run())
vals = [0, 1, 2, 3, 7, 255, 65535, 2**20, 2**27, 2**31 - 1, 2**32 - 1]
fvals = [0.0, -1.0, 1.0, 0.5, 2.0, 1e300, -1e300, float('nan'), float('inf')]
n_err = n_ok = 0
for it in range(int(sys.argv[2])):
    kind = rng.randrange(6)
    if kind == 0 and desc.num_cells:
        c = desc.cells[rng.randrange(desc.num_cells)]; fld = rng.choice(['sensor', 'sub_first', 'sub_count', 'x', 'specularity'])
        if fld == 'sensor': old = c.sensor; c.sensor = rng.choice(vals); rc = run(); c.sensor = old
        elif fld in ('sub_first', 'sub_count'):
            a = getattr(c, fld); k = rng.randrange(3); old = a[k]; a[k] = rng.choice(vals); rc = run(); a[k] = old
        elif fld == 'x':
            k = rng.randrange(3); old = c.x[k]; c.x[k] = rng.choice(fvals); rc = run(); c.x[k] = old
        else: old = c.specularity; c.specularity = rng.choice(fvals); rc = run(); c.specularity = old
    elif kind == 1 and desc.num_subsurfaces:
        s = desc.subsurfaces[rng.randrange(desc.num_subsurfaces)]; fld = rng.choice(['kind', 'target', 'target_edge', 's0', 's1', 't0'])
        old = getattr(s, fld); setattr(s, fld, rng.choice(fvals if fld in ('s0', 's1', 't0') else vals)); rc = run(); setattr(s, fld, old)
    elif kind == 2 and desc.num_emitters:
        e = desc.emitters[rng.randrange(desc.num_emitters)]; fld = rng.choice(['cell', 'edge', 'table', 's_p1', 'start_time', 'duration'])
        old = getattr(e, fld); setattr(e, fld, rng.choice(fvals if fld in ('s_p1', 'start_time', 'duration') else vals)); rc = run(); setattr(e, fld, old)
    elif kind == 3:
        s = desc.sensors[rng.randrange(desc.num_sensors)]; fld = rng.choice(['material', 'base_table', 'scatter_table', 'temperature'])
        old = getattr(s, fld); setattr(s, fld, rng.choice(fvals if fld == 'temperature' else vals)); rc = run(); setattr(s, fld, old)
    elif kind == 4:
        fld = rng.choice(['measurement_steps', 'step_adjustment', 'simulation_time', 'num_materials', 'num_tables'])
        old = getattr(desc, fld)
        new = rng.choice(fvals) if fld == 'simulation_time' else rng.choice([0, 1, 2**31, 2**32 - 1] if fld in ('measurement_steps', 'step_adjustment') else [0])
        setattr(desc, fld, new); rc = run(); setattr(desc, fld, old)
    else:
        src = (psim.Source * 2)()
        src[0].kind = rng.choice([0, 1, 2, 7]); src[0].index = rng.choice(vals); src[0].sign = rng.choice([-1, 0, 1]); src[0].count = rng.choice([0, 1, 5, 2000])
        src[1].kind = 1; src[1].index = 0; src[1].sign = 1; src[1].count = rng.choice([0, 3])
        rc = run(2, src)
    if rc[0]: n_err += 1
    else: n_ok += 1
print("errors", n_err, "ok", n_ok, "final", run())
