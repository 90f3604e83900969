"""The oracle (oracle/model.py + oracle/sim.c, a CPU restatement of the reference) pinned against the golden
fixtures produced by the UNMODIFIED reference (tests/golden/, 16 seeds per case): deterministic known answers to
1e-12, Monte Carlo tallies to 3 sigma."""
import multiprocessing
from concurrent.futures import ProcessPoolExecutor

import numpy as np
import pytest

from oracle.model import OracleModel
from tests import common as T

KAT_CASES = ["linear_demo", "linear_hot_cells", "linear_impurity", "linear_full", "sides_trans", "sige"]


@pytest.mark.parametrize("name", KAT_CASES)
def test_oracle_tables_and_energies_match_reference(name):
    om = OracleModel(T.case_model(name))
    om.prepare()
    kat = T.golden_kat(name)
    assert om.total_energy() == pytest.approx(float(kat["total_energy"]), rel=1e-12)
    init, emit = om.cell_energies()
    np.testing.assert_allclose(om.cell_area, kat["cell_areas"], rtol=1e-12)
    np.testing.assert_allclose(init, kat["cell_init_energy"], rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(emit, kat["cell_emit_energy"], rtol=1e-12, atol=1e-300)
    for mi, mat in enumerate(om.materials):
        arrays = np.stack([mat.freq, mat.vel_la, mat.vel_ta, mat.dens_la, mat.dens_ta])
        np.testing.assert_allclose(arrays[:, ::25], kat["arrays_sub"][mi], rtol=1e-12, atol=1e-300)
        for ti, temp in enumerate(kat["temps"]):
            for kind in range(3):
                tab, total, _ = mat.table(kind, float(temp))
                assert total == pytest.approx(float(kat["sums"][mi][ti][kind]), rel=1e-12)
                np.testing.assert_allclose(tab[::25], kat["tables_sub"][mi][ti][kind], rtol=1e-10, atol=1e-14)
                np.testing.assert_allclose(tab[-1], kat["tables_last"][mi][ti][kind], rtol=1e-10)


def _oracle_one(args):
    name, seed = args
    om = OracleModel(T.case_model(name))
    om.prepare()
    e, f, _, _ = om.run(seed)
    six, temps, fluxes = om.finish_run()
    order = np.argsort(om.sensor_ids, kind="stable")
    return T.run_features(e[order], f[order], om.sim_type, six, temps, fluxes)


def _oracle_features(name, seeds, processes=1):
    if processes == 1:
        return [_oracle_one((name, seed)) for seed in seeds]
    # model set-up is single-threaded Python: one process per seed.  Spawned, not forked: the C restatement runs on OpenMP
    # threads, and a child forked from a process that has already started an OpenMP team hangs in its first parallel region
    with ProcessPoolExecutor(processes, mp_context=multiprocessing.get_context("spawn")) as ex:
        return list(ex.map(_oracle_one, [(name, seed) for seed in seeds]))


@pytest.mark.parametrize("name", ["linear_demo", "linear_full", "sige", "linear_rough", "linear_impurity", "linear_hot_cells"])
def test_oracle_steady_state_parity_with_reference(name):
    """Deviational and full mode, the Si/Ge interface, fully diffuse walls, impurity scattering, cells that start away from
    t_eq: every branch of the restatement against 16 seeds of the unmodified reference."""
    gold = T.golden(name)
    runs = _oracle_features(name, range(1, 9), processes=4)
    T.assert_parity(T.welch_z(runs, gold, "tally_e"), f"oracle {name} energy tallies")
    T.assert_parity(T.welch_z(runs, gold, "tally_f"), f"oracle {name} flux tallies")
    six = T.welch_z(runs, gold, "out6")
    for col in (0, 2, 4):
        T.assert_parity(six[:, col], f"oracle {name} ss column {col}")


@pytest.mark.slow
def test_oracle_transient_trace_parity_with_reference():
    gold = T.golden("sides_trans")
    runs = _oracle_features("sides_trans", range(1, 9), processes=4)
    T.assert_parity(T.welch_z(runs, gold, "tally_e_blk"), "oracle sides_trans energy trace")
    T.assert_parity(T.welch_z(runs, gold, "temp_blk"), "oracle sides_trans temperature trace")
    T.assert_parity(T.welch_z(runs, gold, "flux_blk"), "oracle sides_trans flux trace")


def test_oracle_event_counts_match_survey():
    """Per-phonon work the reference does on linear_demo (145 loop iterations, 64 intervals)."""
    om = OracleModel(T.case_model("linear_demo"))
    om.num_phonons = 50_000
    om.prepare()
    _, _, drift_steps, loop_iters = om.run(3)
    n = om.sources(3)[:, 4].sum()
    assert 60 < drift_steps / n < 69
    assert 135 < loop_iters / n < 155
