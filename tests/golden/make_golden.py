#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Runs in the authoring container only (needs /root/reference and oracle/_ref/psim_ref, built by
`make -C oracle ref`).  For every parity case it
  1. writes the model JSON (psim_b200.configs generators; the kinked wire is read from the reference's
     shipped file because its generator is not part of the hot path),
  2. runs oracle/_ref/psim_ref K times (the reference seeds itself from std::random_device, utils.h:16-21,
     so K invocations = K independent seeds),
  3. stores per-sensor mean and standard deviation ACROSS SEEDS of the raw tallies and of the values the
     reference would print, as tests/golden/<case>.npz, plus the deterministic known answers
     (tables, energies) as tests/golden/<case>.kat.npz.

Nothing in the GPU tests reads /root/reference: they read only the .npz/.json.gz files written here.

usage: python tests/golden/make_golden.py [--seeds 16] [--only case1,case2] [--jobs 8]
"""
from __future__ import annotations

import argparse
import gzip
import json
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from psim_b200 import configs  # noqa: E402
from tests import cases as test_cases  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "psim_ref")
REF_BIN_ITERS3 = os.path.join(ROOT, "oracle", "_ref", "psim_ref_iters3")
REF_JSON = "/root/reference/psim_python/json"
NBLOCKS = 20  # periodic / transient traces are compared as NBLOCKS block means over the recorded steps


def cases():
    """name -> model dict (tests/cases.py); also refreshes the compressed kinked-wire geometry fixture."""
    kinked_path = os.path.join(REF_JSON, "kinked_demo_120_35_spec.json")
    if os.path.exists(kinked_path):
        kinked = json.load(open(kinked_path))
        fixture = os.path.join(HERE, "kinked_demo_120_35_spec.json.gz")
        if not os.path.exists(fixture):
            with gzip.open(fixture, "wt", compresslevel=9) as f:
                json.dump(kinked, f, separators=(",", ":"))
    return test_cases.cases()


def read_run(prefix):
    meta = json.load(open(prefix + ".meta.json"))
    S, R = meta["sensors"], meta["recorded_steps"]
    raw = np.fromfile(prefix + ".bin", dtype=np.uint8)
    off = 0

    def take(dtype, n):
        nonlocal off
        nbytes = np.dtype(dtype).itemsize * n
        a = raw[off:off + nbytes].view(dtype)
        off += nbytes
        return a

    out = {
        "energy": take(np.int32, S * R).reshape(S, R).astype(np.int64),
        "flux": take(np.float64, S * R * 2).reshape(S, R, 2),
        "final_temps": take(np.float64, S * R).reshape(S, R),
        "final_fluxes": take(np.float64, S * R * 2).reshape(S, R, 2),
        "out6": take(np.float64, S * 6).reshape(S, 6),
    }
    assert off == raw.size
    return meta, out


def blocks(a, nb, axis=1):
    """Mean over nb equal blocks along `axis` (trailing remainder dropped)."""
    R = a.shape[axis]
    w = R // nb
    a = np.take(a, np.arange(nb * w), axis=axis)
    shp = list(a.shape)
    shp[axis:axis + 1] = [nb, w]
    return a.reshape(shp).mean(axis=axis + 1)


def features(meta, run):
    """The per-run quantities whose across-seed mean/std become the fixture."""
    S, R = run["energy"].shape
    nb = NBLOCKS if (meta["sim_type"] != 0 and R >= 10 * NBLOCKS) else 1  # steady state: one block
    f = {
        "tally_e": run["energy"].sum(axis=1).astype(np.float64),  # [S]
        "tally_f": run["flux"].sum(axis=1),  # [S,2]
        "out6": run["out6"],  # [S,6]
        "tally_e_blk": blocks(run["energy"].astype(np.float64), nb),  # [S,nb]
        "tally_f_blk": blocks(run["flux"], nb),  # [S,nb,2]
        "temp_blk": blocks(run["final_temps"], nb),  # [S,nb]
        "flux_blk": blocks(run["final_fluxes"], nb),  # [S,nb,2]
        # pooled scalars (one degree of freedom each, known null distribution): total energy and flux tallies, mean exported
        # temperature and fluxes - what a uniform bias over all sensors shows up in
        "pool": np.array([run["energy"].sum(), run["flux"][:, :, 0].sum(), run["flux"][:, :, 1].sum(),
                          run["out6"][:, 0].mean(), run["out6"][:, 2].mean(), run["out6"][:, 4].mean()], dtype=np.float64),
        "total_phonons": np.float64(meta["total_phonons"]),
        "e_post": np.float64(meta["energy_per_phonon_post"]),
        "seconds": np.float64(meta["seconds"]),
    }
    return f


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=16)
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--only", default="")
    ap.add_argument("--iters3", action="store_true",
                    help="run oracle/_ref/psim_ref_iters3 (the reference with MAX_ITERS raised to 3, `make -C oracle ref_iters3`) on "
                         "tests/cases.py:iteration_cases() and write <case>.iters3.npz")
    args = ap.parse_args()
    ref_bin = REF_BIN_ITERS3 if args.iters3 else REF_BIN
    suffix = ".iters3" if args.iters3 else ""
    if not os.path.exists(ref_bin):
        sys.exit(f"{ref_bin} is missing: run `make -C oracle {'ref_iters3' if args.iters3 else 'ref'}` first")
    todo = test_cases.iteration_cases() if args.iters3 else cases()
    if args.only:
        todo = {k: v for k, v in todo.items() if k in args.only.split(",")}
    with tempfile.TemporaryDirectory() as tmp:
        jobs = []
        for name, model in todo.items():
            path = configs.save(model, os.path.join(tmp, name + ".json"))
            if not args.iters3:
                subprocess.run([REF_BIN, "kat", path, os.path.join(tmp, name + ".kat")], check=True,
                               stdout=subprocess.DEVNULL)
            for k in range(args.seeds):
                jobs.append((name, path, os.path.join(tmp, f"{name}.{k}")))

        def run(job):
            name, path, prefix = job
            subprocess.run([ref_bin, "run", path, prefix], check=True, stdout=subprocess.DEVNULL)
            return job

        with ThreadPoolExecutor(args.jobs) as ex:
            for name, _, prefix in ex.map(run, jobs):
                print("done", os.path.basename(prefix), flush=True)

        for name, model in todo.items():
            feats, meta0 = [], None
            for k in range(args.seeds):
                meta, r = read_run(os.path.join(tmp, f"{name}.{k}"))
                meta0 = meta0 or meta
                feats.append(features(meta, r))
            out = {}
            for key in feats[0]:
                stack = np.stack([f[key] for f in feats])
                out[key + "_mean"] = stack.mean(axis=0)
                out[key + "_std"] = stack.std(axis=0, ddof=1).astype(np.float32)
            # calibration of the parity statistic: the same Welch z the GPU test uses, between the two halves
            # of the reference's own seeds (what "agreement" looks like when both sides ARE the reference)
            h = args.seeds // 2
            for key in ("tally_e", "tally_e_blk", "temp_blk"):
                a = np.stack([f[key] for f in feats[:h]])
                b = np.stack([f[key] for f in feats[h:2 * h]])
                se = np.sqrt(a.var(axis=0, ddof=1) / h + b.var(axis=0, ddof=1) / h)
                z = (a.mean(axis=0) - b.mean(axis=0)) / np.where(se > 0, se, np.inf)
                out["selfz_" + key] = np.array([np.abs(z).max(), (np.abs(z) > 3).mean(), z.mean()])
            out["n_seeds"] = np.int64(args.seeds)
            out["num_phonons"] = np.int64(meta0["num_phonons"])
            out["total_energy_pre"] = np.float64(meta0["total_energy_pre"])
            out["energy_per_phonon_pre"] = np.float64(meta0["energy_per_phonon_pre"])
            out["sensor_ids"] = np.array(meta0["sensor_ids"], dtype=np.int64)
            out["sensor_areas"] = np.array(meta0["sensor_areas"], dtype=np.float64)
            out["settings_json"] = np.array(json.dumps(model["settings"]))
            np.savez_compressed(os.path.join(HERE, name + suffix + ".npz"), **out)
            if args.iters3:
                print("wrote", name + suffix, flush=True)
                continue

            # deterministic known answers (tables sub-sampled every 25th bin + full-precision sums)
            kmeta = json.load(open(os.path.join(tmp, name + ".kat.meta.json")))
            kraw = np.fromfile(os.path.join(tmp, name + ".kat.bin"), dtype=np.float64)
            nm, nt = kmeta["num_materials"], len(kmeta["temps"])
            per_mat = 5000 + nt * 6000
            kraw = kraw.reshape(nm, per_mat)
            arrays = kraw[:, :5000].reshape(nm, 5, 1000)
            tables = kraw[:, 5000:].reshape(nm, nt, 3, 1000, 2)
            np.savez_compressed(
                os.path.join(HERE, name + ".kat.npz"),
                total_energy=np.float64(kmeta["total_energy"]), temps=np.array(kmeta["temps"]),
                sums=np.array(kmeta["sums"]),  # [mat][temp][base,emit,scatter]
                arrays_sub=arrays[:, :, ::25].copy(),  # freq, vel_la, vel_ta, dens_la, dens_ta
                tables_sub=tables[:, :, :, ::25, :].copy(), tables_last=tables[:, :, :, -1, :].copy(),
                cell_areas=np.array(kmeta["cell_areas"]), cell_init_energy=np.array(kmeta["cell_init_energy"]),
                cell_emit_energy=np.array(kmeta["cell_emit_energy"]))
            print("wrote", name, flush=True)


if __name__ == "__main__":
    main()
