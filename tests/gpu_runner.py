"""Drives the CUDA path through the C ABI for the GPU tests, smoke() and bench.py."""
from __future__ import annotations

import numpy as np

from psim_b200 import lib as psim
from tests import common as T


def gpu_run_case(model: psim.Model, seed: int, *, shards: int = 1, steps_per_launch: int = 0, device: int = 0,
                 options: dict | None = None, finish: bool = True):
    """One run of `model` on the GPU.  shards > 1 runs the shards one after the other on the same device
    ("virtual shards") and sums their integer tallies exactly as the NCCL all-reduce would."""
    model.prepare()
    desc = model.describe()
    src, n = model.sources(seed)
    e_tot = f_tot = None
    stats = []
    for shard in range(shards):
        g = psim.GpuSimulator(desc, device)
        try:
            g.set_option("steps_per_launch", steps_per_launch)
            for k, v in (options or {}).items():
                g.set_option(k, v)
            g.set_sources(src, n, seed, shard, shards)
            g.run()
            e, f, fx = g.tallies(fixed=True)
            stats.append(g.stats().as_dict())
        finally:
            g.close()
        e_tot = e.astype(np.int64) if e_tot is None else e_tot + e
        f_tot = fx if f_tot is None else f_tot + fx
    energy = e_tot.astype(np.int32)
    flux = f_tot.astype(np.float64) / 256.0
    out = {"energy": energy, "fixed": f_tot, "flux": flux, "stats": stats,
           "sources": [(src[i].kind, src[i].index, src[i].sign, src[i].count) for i in range(n)]}
    if finish:
        model.set_tallies(energy, flux)
        model.finish_run(0)
        six, temps, fluxes = model.results(0)
        out.update(six=six, temps=temps, fluxes=fluxes, e_post=model.energy_per_phonon)
        model.next_run()
    return out


def gpu_features(name: str, seeds, **kw):
    model = T.load_model(T.case_model(name))
    sim_type = model.info.sim_type
    runs = []
    for seed in seeds:
        r = gpu_run_case(model, seed, **kw)
        runs.append(T.run_features(r["energy"], r["flux"], sim_type, r["six"], r["temps"], r["fluxes"]))
    return runs
