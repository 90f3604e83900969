#!/usr/bin/env python
"""Where the time of a whole `psim_model_run` goes beyond its kernels: every shipped model three times in one process with
PSIM_TIMING=1 (host_api.cpp prints its phase times on stderr: prepare / create / set_sources / run / tallies / epilogue /
destroy).  usage (under gpurun): python tools/gpu_phase_times.py [model ...] 2> phases.log"""
import json
import os
import sys
import time

os.environ["PSIM_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psim_b200 import configs  # noqa: E402
from psim_b200 import lib as psim  # noqa: E402
from tests import cases  # noqa: E402


def models():
    m = {
        "linear_demo": configs.linear().to_dict(),
        "sides_ss": configs.linear_sides().to_dict(),
        "sides_per": configs.linear_sides(sim_type=1, step_interval=4).to_dict(),
        "sige": configs.si_ge_grid().to_dict(),
    }
    kinked = cases.kinked_model()
    if kinked is not None:
        m["kinked"] = kinked
    return m


def main():
    want = sys.argv[1:]
    for name, model in models().items():
        if want and name not in want:
            continue
        m = psim.Model(text=json.dumps(model))
        for rep in range(3):
            print(f"== {name} run {rep}", file=sys.stderr, flush=True)
            t0 = time.perf_counter()
            st = m.run(device=0, seed=1 + rep)
            wall = (time.perf_counter() - t0) * 1e3
            print(f"== {name} run {rep}: whole run {wall:.2f} ms, kernels {st.kernel_ms:.2f} ms", file=sys.stderr, flush=True)
        m.close()


if __name__ == "__main__":
    main()
