#!/usr/bin/env python
"""Recorded windows over the lattice image (option "lattice_recorded"): the work-queue kernel's load-balanced tally
(kernels.cuh: tally_lattice) against the lock-step kernel's lane-by-lane runs (device_core.cuh: lattice_runs), bit for bit,
and against the fine-cell run of the same seed.  usage (under gpurun): python tools/gpu_lattice_check.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psim_b200 import configs  # noqa: E402
from tests import common as T  # noqa: E402
from tests.gpu_runner import gpu_run_case  # noqa: E402

cases = {"sides_per": T.load_model(configs.linear_sides(sim_type=1, step_interval=4, num_phonons=60_000).to_dict())}
if "kinked_spec" in T.all_case_names():
    cases["kinked_spec"] = T.load_model(T.case_model("kinked_spec"), num_phonons=100_000)
for name, model in cases.items():
    fine = gpu_run_case(model, 3, options={"lattice_recorded": 0, "tally_shared": 0}, finish=False)
    ref = gpu_run_case(model, 3, options={"kernel": 1, "lattice_recorded": 1, "tally_shared": 0}, finish=False)
    for opts in ({"kernel": 2, "lattice_recorded": 1, "tally_shared": 0}, {"kernel": 0, "lattice_recorded": 1, "tally_shared": 0},
                 {"kernel": 2, "lattice_recorded": 1, "tally_shared": 0, "queue_slots": 64}):
        got = gpu_run_case(model, 3, options=opts, finish=False)
        same = np.array_equal(got["energy"], ref["energy"]) and np.array_equal(got["fixed"], ref["fixed"])
        print(name, opts, "== lock-step:", same, "events", got["stats"][0]["events"], ref["stats"][0]["events"],
              "| energy diff entries", int((got["energy"] != ref["energy"]).sum()), "of", got["energy"].size,
              "sum", int(got["energy"].astype(np.int64).sum()), int(ref["energy"].astype(np.int64).sum()))
    d = np.abs(ref["energy"].astype(np.int64) - fine["energy"]).sum() / max(np.abs(fine["energy"].astype(np.int64)).sum(), 1)
    print(name, "lattice vs fine cells, same seed: |diff| / |tally| =", round(float(d), 5), "events", ref["stats"][0]["events"], fine["stats"][0]["events"])
