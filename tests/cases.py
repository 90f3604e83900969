"""The parity cases: name -> model dict, at the reduced phonon counts the golden fixtures were generated with.

Shared by tests/golden/make_golden.py (which runs the reference on them) and by the tests (which run the CUDA
path, or the CPU-side emulation of its device functions, on the SAME model files).  BASELINE.json's configs:
linear_demo, kinked (specular as shipped + a diffuse variant), linear_sides periodic / transient, full
(non-deviational) mode, and the synthetic Si/Ge grid.
"""
from __future__ import annotations

import gzip
import json
import os

from psim_b200 import configs

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
KINKED_GZ = os.path.join(GOLDEN, "kinked_demo_120_35_spec.json.gz")
REF_KINKED = "/root/reference/psim_python/json/kinked_demo_120_35_spec.json"

HOLLAND_SI = {
    "name": "Silicon",
    "d_data": {"la_data": [-2.01e-07, 9010.0, 0.0], "max_freq_la": 7.63916048e13,
               "ta_data": [-2.26e-07, 5230.0, 0.0], "max_freq_ta": 3.0100793072e13},
    "r_data": {"b_l": 2.0e-24, "b_tn": 9.3e-13, "b_tu": 5.5e-18, "b_i": 1.2e-45, "w": 2.417e13},
}


def kinked_model():
    """The reference's shipped kinked-wire model (its generator is outside the hot path; the geometry is a fixture)."""
    if os.path.exists(KINKED_GZ):
        with gzip.open(KINKED_GZ, "rt") as f:
            return json.load(f)
    if os.path.exists(REF_KINKED):
        return json.load(open(REF_KINKED))
    return None


def cases(include_kinked: bool = True):
    c = {}
    c["linear_demo"] = configs.linear(num_phonons=400_000).to_dict()
    c["linear_diffuse"] = configs.linear(num_phonons=200_000, spec=0.3).to_dict()
    c["linear_rough"] = configs.linear(num_phonons=200_000, spec=0.0).to_dict()
    c["linear_hot_cells"] = configs.linear(num_phonons=200_000, t_init=305.0).to_dict()
    c["linear_impurity"] = configs.linear(num_phonons=200_000, material=HOLLAND_SI).to_dict()
    c["linear_full"] = configs.full_mode(configs.linear(num_phonons=200_000).to_dict(), t_init=25.0,
                                         temp_map={310: 30.0, 290: 20.0})
    c["sides_ss"] = configs.linear_sides(num_phonons=100_000).to_dict()
    c["sides_per"] = configs.linear_sides(num_phonons=100_000, sim_type=1, step_interval=4).to_dict()
    c["sides_trans"] = configs.linear_sides(num_phonons=100_000, sim_type=2, step_interval=4, start_time=0.1,
                                            duration=0.15).to_dict()
    c["sides_per_full"] = configs.full_mode(
        configs.linear_sides(num_phonons=100_000, sim_type=1, step_interval=4).to_dict(), t_init=25.0,
        temp_map={330: 30.0, 270: 20.0, 300.0: 25.0})
    c["sige"] = configs.si_ge_grid(num_phonons=200_000).to_dict()
    if include_kinked:
        kinked = kinked_model()
        if kinked is not None:
            c["kinked_spec"] = configs.with_settings(kinked, num_phonons=50_000)
            c["kinked_diffuse"] = configs.with_specularity(configs.with_settings(kinked, num_phonons=30_000), 0.5)
            c["kinked_rough"] = configs.with_specularity(configs.with_settings(kinked, num_phonons=30_000), 0.0)
    return c


def iteration_cases():
    """Cases for the re-iteration of a run (model.cpp:159-172 with MAX_ITERS raised to 3): steady state (new
    t_eq, per-sensor steady temperatures and tables, cells that now emit), periodic (tables only) and transient (one heat
    capacity and scatter table per sensor and measurement step).  The wall temperatures are far apart so that the sensors
    do move by more than the reference's 0.1 % / 2 % stability thresholds and every case really iterates three times."""
    c = {}
    c["linear_demo"] = configs.linear(num_phonons=150_000, t_high=340, t_low=280).to_dict()
    c["sides_per"] = configs.linear_sides(num_phonons=60_000, sim_type=1, step_interval=4, t_high=360, t_low=250).to_dict()
    c["sides_trans"] = configs.linear_sides(num_phonons=60_000, sim_type=2, step_interval=4, start_time=0.1, duration=0.15,
                                            t_high=360, t_low=250).to_dict()
    return c


def split_bar(num_phonons: int, split=(7, 12)) -> dict:
    """linear_demo's bar with two of its 20 rectangles cut into a lower and an upper half: the full-height edges of
    their neighbours then face TWO cells each (partial transition sub-surfaces, compositeSurface.cpp:47-66).  The
    reference cannot serve as the oracle here - it picks a sub-surface by testing the hit point against the infinite
    line of each (geometry.cpp:88-90), i.e. always the first of two collinear ones - so the check is self-consistency:
    cutting cells must not change the physics."""
    m = configs.ModelFile(num_measurements=1000, sim_time=10, num_phonons=num_phonons, t_eq=300)
    name = m.material(configs.SILICON)
    for i in range(20):
        sid = m.sensor(name, 300.0)
        if i in split:
            m.rectangle((i * 50.0, 0.0), ((i + 1) * 50.0, 100.0), sid, 1)
            m.rectangle((i * 50.0, 100.0), ((i + 1) * 50.0, 200.0), sid, 1)
        else:
            m.rectangle((i * 50.0, 0.0), ((i + 1) * 50.0, 200.0), sid, 1)
    m.emit_surface((0.0, 0.0), (0.0, 200.0), 310)
    m.emit_surface((1000.0, 0.0), (1000.0, 200.0), 290)
    return m.to_dict()
