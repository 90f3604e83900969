#!/usr/bin/env python
"""Run one shipped model at full size (for ncu): python tools/profile_model.py kinked|sides_per|sides_trans|linear|sides_ss [kernel] [phonons]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psim_b200 import configs, lib as psim  # noqa: E402
from tests import cases  # noqa: E402

which = sys.argv[1]
kernel = int(sys.argv[2]) if len(sys.argv) > 2 else -1
model = {"kinked": lambda: cases.kinked_model(),
         "sides_per": lambda: configs.linear_sides(sim_type=1, step_interval=4).to_dict(),
         "sides_trans": lambda: configs.linear_sides(sim_type=2, step_interval=4, start_time=0.1, duration=0.15).to_dict(),
         "linear": lambda: configs.linear().to_dict(), "sides_ss": lambda: configs.linear_sides().to_dict(),
         "sige": lambda: configs.si_ge_grid().to_dict()}[which]()
m = psim.Model(text=json.dumps(model))
if len(sys.argv) > 3:
    m.set_num_phonons(int(float(sys.argv[3])))
m.prepare()
g = psim.GpuSimulator(m.describe(), 0)
if kernel >= 0:
    g.set_option("kernel", kernel)
for kv in os.environ.get("PSIM_OPTS", "").split(","):  # e.g. PSIM_OPTS=queue_slots=64,tally_shared=0
    if "=" in kv:
        g.set_option(kv.split("=")[0], int(kv.split("=")[1]))
for rep in range(2):
    src, n = m.sources(1 + rep)
    g.set_sources(src, n, 1 + rep, 0, 1)
    g.run()
    st = g.stats()
    print(which, "kernel", kernel, "ms", round(st.kernel_ms, 2), "Gds/s", round(st.drift_steps / (st.kernel_ms * 1e-3) / 1e9, 2), "launches", st.launches,
          "events/ds", round(st.events / st.drift_steps, 3), "peak_alive", st.peak_alive)
g.close()
