"""GPU parity tests proper: the CUDA path, called through the C ABI, against the golden fixtures made from 16
seeds of the unmodified reference (tests/golden/make_golden.py).  Statistical bar from BASELINE.json: per-sensor
temperatures / fluxes (steady state) and traces (periodic, transient) within 3 sigma, sigma from >= 8 seeds of
each implementation; integer bookkeeping bit-exact across shard counts."""
import numpy as np
import pytest

from tests import common as T
from tests.gpu_runner import gpu_features, gpu_run_case

pytestmark = pytest.mark.gpu

SEEDS = list(range(1, 9))

STEADY = ["linear_demo", "linear_diffuse", "linear_hot_cells", "linear_impurity", "linear_full", "sides_ss", "sige",
          "kinked_spec", "kinked_diffuse"]
TRACES = ["sides_per", "sides_trans", "sides_per_full"]


@pytest.mark.parametrize("name", STEADY)
def test_steady_state_parity(name):
    if name not in T.all_case_names():
        pytest.skip("fixture geometry missing")
    gold = T.golden(name)
    runs = gpu_features(name, SEEDS)
    T.assert_parity(T.welch_z(runs, gold, "tally_e"), f"{name} energy tallies")
    T.assert_parity(T.welch_z(runs, gold, "tally_f"), f"{name} flux tallies")
    six = T.welch_z(runs, gold, "out6")
    T.assert_parity(six[:, 0], f"{name} temperature column")   # T as written to ss_*.txt
    T.assert_parity(six[:, 2], f"{name} x-flux column")
    T.assert_parity(six[:, 4], f"{name} y-flux column")


@pytest.mark.parametrize("name", TRACES)
def test_trace_parity(name):
    gold = T.golden(name)
    runs = gpu_features(name, SEEDS)
    T.assert_parity(T.welch_z(runs, gold, "tally_e_blk"), f"{name} energy trace")
    T.assert_parity(T.welch_z(runs, gold, "tally_f_blk"), f"{name} flux trace")
    T.assert_parity(T.welch_z(runs, gold, "temp_blk"), f"{name} temperature trace")
    T.assert_parity(T.welch_z(runs, gold, "flux_blk"), f"{name} exported flux trace")


@pytest.mark.parametrize("name", ["linear_demo", "sides_trans", "sige"])
def test_integer_bookkeeping_is_shard_invariant(name):
    model = T.load_model(T.case_model(name), num_phonons=60_000)
    ref = gpu_run_case(model, 7, shards=1, finish=False)
    for shards in (2, 4, 8):
        got = gpu_run_case(model, 7, shards=shards, finish=False)
        assert got["sources"] == ref["sources"]                      # emitted phonon counts
        assert np.array_equal(got["energy"], ref["energy"])          # int32 energy tallies, bit for bit
        assert np.array_equal(got["fixed"], ref["fixed"])            # fixed-point flux tallies, bit for bit
        assert sum(s["drift_steps"] for s in got["stats"]) == ref["stats"][0]["drift_steps"]


@pytest.mark.parametrize("name", ["linear_demo", "sides_per"])
def test_steps_per_launch_does_not_change_results(name):
    model = T.load_model(T.case_model(name), num_phonons=50_000)
    ref = gpu_run_case(model, 3, steps_per_launch=1, finish=False)
    for spl in (2, 7):
        got = gpu_run_case(model, 3, steps_per_launch=spl, finish=False)
        assert np.array_equal(got["energy"], ref["energy"])
        assert np.array_equal(got["fixed"], ref["fixed"])


def test_tally_paths_agree():
    model = T.load_model(T.case_model("sides_per"), num_phonons=50_000)
    ref = gpu_run_case(model, 5, options={"tally_shared": 0, "tally_aggregate": 0}, finish=False)
    for opts in ({"tally_shared": 1, "tally_aggregate": 0}, {"tally_shared": 1, "tally_aggregate": 1},
                 {"tally_shared": 0, "tally_aggregate": 1}):
        got = gpu_run_case(model, 5, options=opts, finish=False)
        assert np.array_equal(got["energy"], ref["energy"]), opts
        assert np.array_equal(got["fixed"], ref["fixed"]), opts


def test_cell_population_matches_emulation():
    """Phonons per cell after a few measurement steps: GPU vs the host build of the same device functions.
    The two differ only in floating-point library rounding, so the populations agree statistically, and the
    totals emitted so far agree exactly."""
    from psim_b200 import lib as psim
    model = T.load_model(T.case_model("linear_demo"), num_phonons=50_000)
    model.prepare()
    desc = model.describe()
    src, n = model.sources(11)
    g = psim.GpuSimulator(desc, 0)
    g.set_sources(src, n, 11, 0, 1)
    g.run_steps(0, 40)
    hist = g.cell_histogram()
    alive = g.alive()
    g.close()
    assert int(hist.sum()) == alive
    emu = T.emu_run(model, 11, want_alive=True, want_hist=True)
    assert abs(int(emu["alive"][39]) - alive) <= 6 * np.sqrt(alive)
    assert np.abs(hist.astype(float) - emu["hist"][39].astype(float)).max() <= 6 * np.sqrt(hist.max() + 1)
