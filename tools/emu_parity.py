#!/usr/bin/env python
"""CPU-side check of the device algorithm (tests/emu, the host build of device_core.cuh) against the golden
fixtures: 8 seeds per case, Welch z per sensor / block.  Usage: python tools/emu_parity.py [case ...]"""
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import common as T  # noqa: E402


def one(args):
    name, seed = args
    model = T.load_model(T.case_model(name))
    model.prepare()
    sim_type = model.info.sim_type
    r = T.emu_run(model, seed, steps_per_pass=int(os.environ.get("SPP", "1")))
    model.set_tallies(r["energy"], r["flux"])
    model.finish_run(0)
    six, temps, fluxes = model.results(0)
    f = T.run_features(r["energy"], r["flux"], sim_type, six, temps, fluxes)
    f["_steps"] = r["drift_steps"]
    f["_e_post"] = model.energy_per_phonon
    return name, f


def main():
    names = sys.argv[1:] or T.all_case_names()
    T.emu_lib()
    seeds = range(1, 9)
    t0 = time.time()
    with ProcessPoolExecutor(8) as ex:
        res = list(ex.map(one, [(n, s) for n in names for s in seeds]))
    print(f"emulation took {time.time() - t0:.1f} s")
    for name in names:
        runs = [f for n, f in res if n == name]
        gold = T.golden(name)
        print(f"== {name}: drift-steps/phonon {np.mean([r['_steps'] for r in runs]) / int(gold['num_phonons']):.1f}"
              f"  e_post ours {np.mean([r['_e_post'] for r in runs]):.6g} ref {float(gold['e_post_mean']):.6g}")
        for key in ("tally_e", "tally_f", "out6", "tally_e_blk", "temp_blk", "flux_blk"):
            z = T.welch_z(runs, gold, key)
            if key == "out6":
                for col, lab in ((0, "T"), (2, "qx"), (4, "qy")):
                    print(f"   out6.{lab:3s}", T.parity_summary(z[:, col]))
            else:
                print(f"   {key:11s}", T.parity_summary(z))


if __name__ == "__main__":
    main()
