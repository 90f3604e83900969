"""Model-file generators for the psim JSON input schema.

The schema is the one the reference loader reads (reference psim/src/inputManager.cpp:17-102):
``settings`` / ``materials`` / ``sensors`` / ``cells`` / ``emit_surfaces``.  The generators here are
written from scratch; ``tests/test_configs.py`` checks that ``linear_demo()`` and ``linear_sides()``
reproduce the shipped reference files number for number when the reference tree is mounted.

Units follow the reference: nm, ns, K (psim_python/psim/builder_tools.py:320-327).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence, Tuple

# Material constants (reference psim_python/psim/pre_builts.py:9-28).  The shipped JSON files use the
# "Jean 2014" silicon set; the synthetic Si/Ge grid uses that and the germanium set.
SILICON = {
    "name": "Silicon",
    "d_data": {"la_data": [-2.22e-07, 9260.0, 0.0], "max_freq_la": 7.63916048e13,
               "ta_data": [-2.28e-07, 5240.0, 0.0], "max_freq_ta": 3.0100793072e13},
    "r_data": {"b_l": 1.3e-24, "b_tn": 9e-13, "b_tu": 1.9e-18, "b_i": 0, "w": 2.42e13},
}
GERMANIUM = {
    "name": "Germanium",
    "d_data": {"la_data": [-1.50e-07, 5630.0, 0.0], "max_freq_la": 4.45236386e13,
               "ta_data": [-1.13e-07, 2600.0, 0.0], "max_freq_ta": 1.4937724175e13},
    "r_data": {"b_l": 2.3e-24, "b_tn": 30.0e-13, "b_tu": 1.5e-18, "b_i": 0.0, "w": 1.23e13},
}

Point = Tuple[float, float]


class ModelFile:
    """Accumulates sensors, triangular cells and emitting surfaces, then serialises to the psim schema."""

    def __init__(self, *, num_measurements: int, sim_time: float, num_phonons: int, t_eq: float,
                 sim_type: int = 0, step_interval: int = 0, phasor_sim: bool = False,
                 num_runs: Optional[int] = None):
        self.settings: Dict = {
            "num_measurements": num_measurements, "sim_time": sim_time, "num_phonons": num_phonons,
            "t_eq": t_eq, "sim_type": sim_type, "step_interval": step_interval, "phasor_sim": phasor_sim,
        }
        if num_runs is not None:
            self.settings["num_runs"] = num_runs
        self.materials: List[Dict] = []
        self.sensors: List[Dict] = []
        self.cells: List[Dict] = []
        self.surfaces: List[Dict] = []

    def material(self, mat: Dict) -> str:
        if all(m["name"] != mat["name"] for m in self.materials):
            self.materials.append(json.loads(json.dumps(mat)))
        return mat["name"]

    def sensor(self, material: str, t_init: float) -> int:
        sid = len(self.sensors)
        self.sensors.append({"id": sid, "material": material, "t_init": t_init})
        return sid

    def triangle(self, p1: Point, p2: Point, p3: Point, sensor_id: int, spec: float) -> None:
        self.cells.append({
            "triangle": {"p1": {"x": p1[0], "y": p1[1]}, "p2": {"x": p2[0], "y": p2[1]},
                         "p3": {"x": p3[0], "y": p3[1]}},
            "sensorID": sensor_id, "specularity": spec})

    def rectangle(self, lower_left: Point, upper_right: Point, sensor_id: int, spec: float) -> None:
        # two clockwise right triangles sharing the anti-diagonal, the convention every shipped model uses
        # (builder_tools.py:403-413) and that the reference loader treats as "clockwise" (cell.cpp:120)
        a, b = lower_left, upper_right
        self.triangle(a, (a[0], b[1]), (b[0], a[1]), sensor_id, spec)
        self.triangle(b, (b[0], a[1]), (a[0], b[1]), sensor_id, spec)

    def emit_surface(self, p1: Point, p2: Point, temp: float, duration: float = 0.0,
                     start_time: float = 0.0) -> None:
        sq_len = (p2[0] - p1[0]) ** 2 + (p2[1] - p1[1]) ** 2
        self.surfaces.append({"p1": {"x": p1[0], "y": p1[1]}, "p2": {"x": p2[0], "y": p2[1]}, "temp": temp,
                              "duration": duration, "start_time": start_time, "length": sq_len})

    def to_dict(self) -> Dict:
        sim_time = self.settings["sim_time"]
        surfaces = []
        for s in self.surfaces:
            s = dict(s)
            if s["duration"] == 0.0:  # "always on" is written as the whole run (builder_tools.py:536-545)
                s["duration"] = sim_time
            surfaces.append(s)
        surfaces.sort(key=lambda s: -s["length"])  # stable, like the reference exporter
        return {"settings": dict(self.settings), "materials": self.materials, "sensors": self.sensors,
                "cells": self.cells, "emit_surfaces": surfaces}

    def write(self, path: str) -> str:
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with open(path, "w", encoding="utf-8") as f:
            json.dump(self.to_dict(), f)
        return path


def linear(*, num_cells: int = 20, t_high: float = 310, t_low: float = 290, t_init: float = 300.0,
           t_eq: float = 300, x_base: float = 50, y_base: float = 200, spec: float = 1,
           sim_time: float = 10, num_measurements: int = 1000, num_phonons: int = 5_000_000,
           sim_type: int = 0, step_interval: int = 0, material: Dict = SILICON,
           num_runs: Optional[int] = None) -> ModelFile:
    """A bar of ``num_cells`` rectangles, hot wall on the left, cold wall on the right.

    Defaults give the shipped ``linear_demo.json`` (pre_builts.py:34-58 with linear_demo.py's arguments).
    """
    m = ModelFile(num_measurements=num_measurements, sim_time=sim_time, num_phonons=num_phonons, t_eq=t_eq,
                  sim_type=sim_type, step_interval=step_interval, num_runs=num_runs)
    name = m.material(material)
    for i in range(num_cells):
        sid = m.sensor(name, t_init)
        m.rectangle((i * x_base, 0.0), ((i + 1) * x_base, y_base), sid, spec)
    m.emit_surface((0.0, 0.0), (0.0, y_base), t_high)
    m.emit_surface((num_cells * x_base, 0.0), (num_cells * x_base, y_base), t_low)
    return m


def linear_sides(*, num_cells: int = 100, t_high: float = 330, t_low: float = 270, t_eq: float = 300.0,
                 x_base: float = 10, y: float = 100, spec: float = 1, sim_time: float = 0.5,
                 num_measurements: int = 1000, num_phonons: int = 10_000_000, sim_type: int = 0,
                 step_interval: int = 0, start_time: float = 0.0, duration: float = 0.0,
                 t_init: Optional[float] = None) -> ModelFile:
    """A wide strip whose top/bottom edges carry hot and cold patches and whose ends sit at the mean.

    Defaults give ``linear_sides_demo_ss.json``; ``sim_type=1, step_interval=4`` gives the periodic file and
    ``sim_type=2, step_interval=4, start_time=0.1, duration=0.15`` the transient one (pre_builts.py:60-138).
    """
    m = ModelFile(num_measurements=num_measurements, sim_time=sim_time, num_phonons=num_phonons, t_eq=t_eq,
                  sim_type=sim_type, step_interval=step_interval)
    name = m.material(SILICON)
    avg = (t_high + t_low) / 2 if t_init is None else t_init
    length = num_cells * x_base
    n_patch = int(length * 0.1 / x_base)
    ny = int(y / x_base)
    for i in range(num_cells):
        for j in range(ny):
            sid = m.sensor(name, avg)
            m.rectangle((i * x_base, j * x_base), ((i + 1) * x_base, (j + 1) * x_base), sid, spec)
    xl, xr = length * 0.3, length * 0.6
    for _ in range(n_patch):
        if duration != 0:
            m.emit_surface((xl, y), (xl + x_base, y), t_high, duration, start_time + duration)
            m.emit_surface((xl, 0.0), (xl + x_base, 0), t_high, duration, start_time)
            m.emit_surface((xr, y), (xr + x_base, y), t_low, duration, start_time)
            m.emit_surface((xr, 0.0), (xr + x_base, 0.0), t_low, duration, start_time + duration)
        else:
            m.emit_surface((xl, y), (xl + x_base, y), t_high)
            m.emit_surface((xl, 0.0), (xl + x_base, 0), t_high)
            m.emit_surface((xr, y), (xr + x_base, y), t_low)
            m.emit_surface((xr, 0.0), (xr + x_base, 0.0), t_low)
        xl += x_base
        xr += x_base
    for j in range(ny):
        m.emit_surface((0.0, j * x_base), (0.0, (j + 1) * x_base), avg)
        m.emit_surface((num_cells * x_base, j * x_base), (num_cells * x_base, (j + 1) * x_base), avg)
    return m


def si_ge_grid(*, nx: int = 10, ny: int = 5, cell: float = 20.0, t_high: float = 310.0, t_low: float = 290.0,
               t_eq: float = 300.0, spec: float = 0.5, sim_time: float = 1.0, num_measurements: int = 1000,
               num_phonons: int = 100_000_000) -> ModelFile:
    """BASELINE.json's synthetic multi-cell Si/Ge structure.

    ``nx`` x ``ny`` squares (2*nx*ny triangular cells, one sensor per square); the left half is silicon, the
    right half germanium; hot wall at x=0, cold wall at x=nx*cell, partially diffuse top and bottom.
    All silicon cells are listed before any germanium cell: the reference orients the back-scatter normal
    of a material interface from the OLDER cell (cell.cpp:85-96,126-132), so only this order keeps its
    phonons inside the mesh; with it, both implementations describe the same physics.
    """
    m = ModelFile(num_measurements=num_measurements, sim_time=sim_time, num_phonons=num_phonons, t_eq=t_eq)
    si = m.material(SILICON)
    ge = m.material(GERMANIUM)
    half = nx // 2
    for mat, cols in ((si, range(0, half)), (ge, range(half, nx))):
        for i in cols:
            for j in range(ny):
                sid = m.sensor(mat, t_eq)
                m.rectangle((i * cell, j * cell), ((i + 1) * cell, (j + 1) * cell), sid, spec)
    for j in range(ny):
        m.emit_surface((0.0, j * cell), (0.0, (j + 1) * cell), t_high)
    for j in range(ny):
        m.emit_surface((nx * cell, j * cell), (nx * cell, (j + 1) * cell), t_low)
    return m


def with_settings(model: Dict, **overrides) -> Dict:
    """Copy of a model dict with some ``settings`` keys replaced (e.g. a reduced ``num_phonons``)."""
    out = dict(model)
    out["settings"] = dict(model["settings"])
    out["settings"].update(overrides)
    return out


def with_specularity(model: Dict, spec: float) -> Dict:
    out = dict(model)
    out["cells"] = [dict(c, specularity=spec) for c in model["cells"]]
    return out


def full_mode(model: Dict, *, t_init: float, temp_map: Dict[float, float]) -> Dict:
    """Non-deviational variant: ``t_eq = 0``, every sensor at ``t_init``, wall temperatures remapped."""
    out = with_settings(model, t_eq=0)
    out["sensors"] = [dict(s, t_init=t_init) for s in model["sensors"]]
    out["emit_surfaces"] = [dict(s, temp=temp_map.get(s["temp"], s["temp"])) for s in model["emit_surfaces"]]
    return out


def save(model: Dict, path: str) -> str:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w", encoding="utf-8") as f:
        json.dump(model, f)
    return path
