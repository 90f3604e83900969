#!/usr/bin/env python
"""Per-window kernel times of one job of the bench workload, for different ways of cutting the measurement steps into
library calls (diagnostic; prints one JSON line per variant)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psim_b200 import configs  # noqa: E402
from psim_b200 import lib as psim  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    model = psim.Model(text=json.dumps(configs.si_ge_grid(num_phonons=n).to_dict()))
    model.prepare()
    info = model.info
    M, R = info.measurement_steps, info.recorded_steps
    first = M - R
    stream = torch.cuda.Stream()
    g = psim.GpuSimulator(model.describe(), 0)

    def job(cuts, seed, label):
        src, cnt = model.sources(seed)
        g.set_sources(src, cnt, seed, 0, 1)
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(cuts))]
        evs[0].record(stream)
        for i in range(len(cuts) - 1):
            g.run_steps(cuts[i], cuts[i + 1], stream.cuda_stream)
            evs[i + 1].record(stream)
        torch.cuda.synchronize()
        ms = [round(evs[i].elapsed_time(evs[i + 1]), 2) for i in range(len(cuts) - 1)]
        st = g.stats()
        print(json.dumps({"variant": label, "cuts": cuts, "ms": ms, "total_ms": round(sum(ms), 2), "kernel_ms": round(st.kernel_ms, 2),
                          "launches": st.launches, "drift_steps": st.drift_steps}), flush=True)

    for rep in range(2):
        job([0, M], 1 + rep, "one call")
        job([0, first - 1] + list(range(first - 1 + 48, M - 1, 48)) + [M - 1], 1 + rep, "bench cuts (48)")
        job([0, first - 1, M - 1], 1 + rep, "two calls")
        job([0, first - 1, M], 1 + rep, "two calls to M")
    g.close()


if __name__ == "__main__":
    main()
