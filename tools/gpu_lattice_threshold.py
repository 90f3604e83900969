#!/usr/bin/env python
"""Where do recorded windows over the lattice image start to pay?  linear_sides periodic (1000 sensors) with longer and longer
measurement steps: fine cells crossed per step at the largest group velocity (the library's `lattice_cells_per_step`) against
kernel ms with "lattice_recorded" 0 and 1.  usage (under gpurun): python tools/gpu_lattice_threshold.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psim_b200 import configs, lib as psim  # noqa: E402

for sim_time in (0.5, 1.0, 2.0, 3.0, 4.0):
    model = configs.linear_sides(sim_type=1, step_interval=4, sim_time=sim_time, num_phonons=4_000_000).to_dict()
    out = {}
    for lr in (0, 1):
        m = psim.Model(text=json.dumps(model))
        m.prepare()
        g = psim.GpuSimulator(m.describe(), 0)
        g.set_option("lattice_recorded", lr)
        best = None
        for rep in range(3):
            src, n = m.sources(1 + rep)
            g.set_sources(src, n, 1 + rep, 0, 1)
            g.run()
            st = g.stats()
            best = st.kernel_ms if best is None or st.kernel_ms < best else best
        out[lr] = (round(best, 2), round(st.events / st.drift_steps, 3))
        g.close()
        m.close()
    print("step", sim_time, "ps: cells per step", round(9260.0 * sim_time / 1000 / 10.0, 2), "fine", out[0], "lattice", out[1])
