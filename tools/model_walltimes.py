#!/usr/bin/env python
"""Wall-clock per model (BASELINE.json: "wall-clock per model"): the reference's shipped configurations at their FULL
phonon counts, end to end through the host API (load JSON -> set-up -> GPU run -> epilogue), plus the kernel-only time
and the drift-step throughput.  One JSON line per model."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psim_b200 import configs  # noqa: E402
from psim_b200 import lib as psim  # noqa: E402
from tests import cases  # noqa: E402


def models():
    m = {
        "linear_demo (5e6, SS)": configs.linear().to_dict(),
        "linear_sides_demo_ss (1e7, SS)": configs.linear_sides().to_dict(),
        "linear_sides_demo_per (1e7, periodic)": configs.linear_sides(sim_type=1, step_interval=4).to_dict(),
        "linear_sides_demo_trans (1e7, transient)": configs.linear_sides(sim_type=2, step_interval=4, start_time=0.1, duration=0.15).to_dict(),
        "linear_demo full mode 25 K (5e6, SS)": configs.full_mode(configs.linear().to_dict(), t_init=25.0, temp_map={310: 30.0, 290: 20.0}),
        "synthetic Si/Ge 100 cells (1e8, SS)": configs.si_ge_grid().to_dict(),
    }
    kinked = cases.kinked_model()
    if kinked is not None:
        m["kinked_demo_120_35_spec (2e7, SS)"] = kinked
        m["kinked_demo_120_35 specularity 0.5 (2e7, SS)"] = configs.with_specularity(kinked, 0.5)
    return m


def main():
    with tempfile.TemporaryDirectory() as tmp:
        for name, model in models().items():
            path = configs.save(model, os.path.join(tmp, "m.json"))
            t0 = time.perf_counter()
            m = psim.Model(path)
            t1 = time.perf_counter()
            st = m.run(device=0, seed=1)
            t2 = time.perf_counter()
            m.export(path, t2 - t1)
            t3 = time.perf_counter()
            six, _, _ = m.results(0, traces=False)
            print(json.dumps({"model": name, "phonons": st.total_phonons, "load_s": round(t1 - t0, 3), "run_s": round(t2 - t1, 3),
                              "export_s": round(t3 - t2, 3), "kernel_ms": round(st.kernel_ms, 2), "launches": st.launches,
                              "steps_per_launch": st.steps_per_launch, "drift_steps": st.drift_steps,
                              "drift_steps_per_s_kernel": st.drift_steps / (st.kernel_ms * 1e-3), "peak_alive": st.peak_alive,
                              "T_first_last": [round(float(six[0, 0]), 3), round(float(six[-1, 0]), 3)]}), flush=True)
            m.close()


if __name__ == "__main__":
    main()
