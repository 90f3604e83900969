"""TEST HELPER (not part of the product package): readers of the result tables psim writes (`ss_<stem>.txt`, `per_<stem>.txt`), used to check the exporter.

The formats are the reference's (psim/src/outputManager.cpp:72-114); its Python tools parse them line by line in
psim_python/psim/plotting_tools.py:84-157 (`parse_ss_data`, `parse_avg_flux`, `parse_periodic_data`).  These readers
return numpy arrays with the same content, so post-processing written against the reference's parsers carries over:

    ss  file:  title line, then per sensor (ascending id)   T  std(T)  qx  std(qx)  qy  std(qy)
    per file:  title line, then per block of `step_interval` measurement steps
                   <centre step of the block>
                   <number of sensors>
                   per sensor (ascending id)   T  qx  qy     (block averages)
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import Optional

import numpy as np

# result files of older reference versions (psim_python/json/results/*.txt, 2022) end at "[s]"
_TITLE = re.compile(r'^(Steady State|Periodic) Results from "(.*)" @ (.*) - Time Taken (\S+)\[s\](?: over (\d+) runs)?$')


@dataclass
class Title:
    kind: str          # "Steady State" or "Periodic"
    model_file: str
    when: str
    seconds: float
    runs: int

    @staticmethod
    def parse(line: str) -> Optional["Title"]:
        m = _TITLE.match(line.strip())
        if not m:
            return None
        return Title(m.group(1), m.group(2), m.group(3), float(m.group(4)), int(m.group(5) or 1))


@dataclass
class SteadyState:
    title: str
    header: Optional[Title]
    table: np.ndarray  # [sensors][6]

    temps = property(lambda self: self.table[:, 0])
    temps_std = property(lambda self: self.table[:, 1])
    x_flux = property(lambda self: self.table[:, 2])
    x_flux_std = property(lambda self: self.table[:, 3])
    y_flux = property(lambda self: self.table[:, 4])
    y_flux_std = property(lambda self: self.table[:, 5])

    def average_flux(self):
        """Mean and sample standard deviation over the sensors of both flux components (parse_avg_flux)."""
        return (float(self.x_flux.mean()), float(self.x_flux.std(ddof=1)), float(self.y_flux.mean()),
                float(self.y_flux.std(ddof=1)))


@dataclass
class Periodic:
    title: str
    header: Optional[Title]
    measurement_steps: np.ndarray  # [blocks] centre step of every block
    temps: np.ndarray              # [blocks][sensors]
    x_flux: np.ndarray             # [blocks][sensors]
    y_flux: np.ndarray             # [blocks][sensors]


def parse_steady_state(text: str) -> SteadyState:
    lines = text.splitlines()
    if not lines:
        raise ValueError("empty steady-state table")
    rows = [ln.split() for ln in lines[1:] if ln.strip()]
    if any(len(r) != 6 for r in rows):
        raise ValueError("a steady-state row needs six columns: T std qx std qy std")
    table = np.array(rows, dtype=np.float64).reshape(len(rows), 6)
    return SteadyState(lines[0], Title.parse(lines[0]), table)


def parse_periodic(text: str) -> Periodic:
    lines = text.splitlines()
    if not lines:
        raise ValueError("empty periodic table")
    body = lines[1:]
    steps, temps, fx, fy = [], [], [], []
    i = 0
    while i < len(body):
        if not body[i].strip():
            i += 1
            continue
        if i + 1 >= len(body):
            raise ValueError("truncated block header")
        step, n = int(body[i]), int(body[i + 1])
        rows = [ln.split() for ln in body[i + 2:i + 2 + n]]
        if len(rows) != n or any(len(r) != 3 for r in rows):
            raise ValueError(f"block at step {step}: expected {n} rows of T qx qy")
        block = np.array(rows, dtype=np.float64).reshape(n, 3)
        steps.append(step)
        temps.append(block[:, 0])
        fx.append(block[:, 1])
        fy.append(block[:, 2])
        i += n + 2
    shape = (len(steps), len(temps[0]) if temps else 0)
    return Periodic(lines[0], Title.parse(lines[0]), np.array(steps, dtype=np.int64), np.array(temps).reshape(shape),
                    np.array(fx).reshape(shape), np.array(fy).reshape(shape))


def read_steady_state(path: str) -> SteadyState:
    with open(path, "r", encoding="utf-8") as f:
        return parse_steady_state(f.read())


def read_periodic(path: str) -> Periodic:
    with open(path, "r", encoding="utf-8") as f:
        return parse_periodic(f.read())
