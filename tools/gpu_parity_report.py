#!/usr/bin/env python
"""The numbers behind the green parity tests, written down: every parity case of tests/test_gpu_parity.py run on the GPU with
the tests' seeds, and for each compared quantity the summary the assertions bound (entries, largest |z|, fraction beyond 3
sigma, mean z, rms z) plus the pooled one-degree-of-freedom z values.  One JSON line per case.
usage (under gpurun): python tools/gpu_parity_report.py > gpurun_out/parity.jsonl"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import common as T  # noqa: E402
from tests.gpu_runner import gpu_features, gpu_run_case  # noqa: E402
from tests.test_gpu_parity import HIGH_STATISTICS, SEEDS, STEADY, TRACES  # noqa: E402


def summary(z):
    s = T.parity_summary(np.asarray(z))
    return {k: (round(v, 3) if isinstance(v, float) else v) for k, v in s.items()}


for name in STEADY + TRACES:
    if name not in T.all_case_names():
        continue
    gold = T.golden(name)
    factor = HIGH_STATISTICS.get(name, 1)
    steady = name in STEADY
    if factor > 1:
        model = T.load_model(T.case_model(name), num_phonons=factor * T.case_model(name)["settings"]["num_phonons"])
        runs = []
        for seed in SEEDS:
            r = gpu_run_case(model, seed)
            runs.append(T.run_features(r["energy"], r["flux"], 0, r["six"], r["temps"], r["fluxes"]))
    else:
        runs = gpu_features(name, range(1, 33) if name == "linear_full" else SEEDS)
    rec = {"case": name, "gpu_seeds": len(runs), "reference_seeds": int(gold["n_seeds"]), "phonon_factor": factor}
    if steady:
        rec["energy_tallies"] = summary(T.welch_z(runs, gold, "tally_e", 1.0 / factor))
        rec["flux_tallies"] = summary(T.welch_z(runs, gold, "tally_f", 1.0 / factor))
        six = T.welch_z(runs, gold, "out6")
        rec["temperature_column"], rec["x_flux_column"], rec["y_flux_column"] = summary(six[:, 0]), summary(six[:, 2]), summary(six[:, 4])
        pooled = T.assert_pooled(runs, gold, name, 1.0 / factor, limit=1e9)
        rec["pooled_z_E_Fx_Fy_T_qx_qy"] = None if pooled is None else [round(float(x), 2) for x in pooled]
    else:
        for key, label in (("tally_e_blk", "energy_trace"), ("tally_f_blk", "flux_trace"), ("temp_blk", "temperature_trace"), ("flux_blk", "exported_flux_trace")):
            rec[label] = summary(T.welch_z(runs, gold, key))
    print(json.dumps(rec), flush=True)
