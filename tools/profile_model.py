#!/usr/bin/env python
"""Run one shipped model at full size (for ncu): python tools/profile_model.py kinked|sides_per|linear"""
import os, sys, json, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psim_b200 import configs, lib as psim
from tests import cases
which = sys.argv[1]
model = {"kinked": lambda: cases.kinked_model(), "sides_per": lambda: configs.linear_sides(sim_type=1, step_interval=4).to_dict(),
         "linear": lambda: configs.linear().to_dict(), "sides_ss": lambda: configs.linear_sides().to_dict()}[which]()
m = psim.Model(text=json.dumps(model))
st = m.run(device=0, seed=1)
print(which, st.kernel_ms, st.drift_steps / (st.kernel_ms * 1e-3) / 1e9, st.launches)
