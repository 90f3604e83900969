#!/usr/bin/env python
"""A/B of build variants (psim_b200/build.py:build_variant) on one GPU: kernel time of one job of the bench workload
and of the shipped models that stress other parts of the kernel.  One process per variant (the library path is fixed
at import); prints one JSON line per (variant, model).

usage:  python tools/gpu_ab_variants.py build            (here, no GPU: compiles the variants)
        python tools/gpu_ab_variants.py run [name ...]   (under gpurun)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {  # round 2: more resident warps with fewer registers; what the flight loop's fast path takes on; the pass-fusion threshold
    "b1024_q96": {"PSIM_BLOCK": 1024, "PSIM_QUEUE_SLOTS": 96},
    "b896_q104": {"PSIM_BLOCK": 896, "PSIM_QUEUE_SLOTS": 104},
    "b896_q96": {"PSIM_BLOCK": 896, "PSIM_QUEUE_SLOTS": 96},
    "fast00": {"PSIM_FAST_WALLS": 0, "PSIM_FAST_COMPOSITE": 0},
    "fast11": {"PSIM_FAST_WALLS": 1, "PSIM_FAST_COMPOSITE": 1},
    "fuse16": {"PSIM_FUSE_LANES": 16},
    "fuse33": {"PSIM_FUSE_LANES": 33},
}


def variant_path(name):
    return os.path.join(ROOT, "psim_b200", "lib", "variants", f"libpsim_b200_{name}.so")


def child(label):
    import time

    from psim_b200 import configs
    from psim_b200 import lib as psim
    from tests import cases

    models = {
        "sige_1e8": configs.si_ge_grid().to_dict(),
        "sides_per_1e7": configs.linear_sides(sim_type=1, step_interval=4).to_dict(),
        "linear_5e6": configs.linear().to_dict(),
    }
    kinked = cases.kinked_model()
    if kinked is not None:
        models["kinked_2e7"] = kinked
    for name, model in models.items():
        m = psim.Model(text=json.dumps(model))
        best = None
        for rep in range(3):  # the first run of a process pays context creation and allocation
            t0 = time.perf_counter()
            st = m.run(device=0, seed=1 + rep)
            wall = time.perf_counter() - t0
            if rep and (best is None or st.kernel_ms < best["kernel_ms"]):
                best = {"variant": label, "model": name, "kernel_ms": round(st.kernel_ms, 2), "wall_ms": round(wall * 1e3, 1),
                        "launches": st.launches, "steps_per_launch": st.steps_per_launch, "warps": st.warps,
                        "drift_steps_per_s": st.drift_steps / (st.kernel_ms * 1e-3), "kernel": st.kernel}
        print(json.dumps(best), flush=True)
        m.close()


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        from psim_b200 import build
        for name, defines in VARIANTS.items():
            print(build.build_variant(name, defines))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(sys.argv[2])
        return
    names = sys.argv[2:] or ["default", *VARIANTS]
    for name in names:
        env = dict(os.environ)
        if name != "default":
            env["PSIM_B200_LIB"] = variant_path(name)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child", name], env=env, check=False)


if __name__ == "__main__":
    main()
