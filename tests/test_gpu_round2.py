"""GPU tests added in round 2: phasor mode on the device, phonons-per-cell across shards, the Philox block budget of the
packed kernels, the many-sensor tally path against the staged one, and the compressed cell records."""
import numpy as np
import pytest

from psim_b200 import configs
from psim_b200 import lib as psim
from tests import common as T
from tests.gpu_runner import gpu_run_case

pytestmark = pytest.mark.gpu


def test_phasor_mode_is_exact_kinematics_on_the_device():
    """phasor_sim (phononBuilder.cpp:42-49, modelSimulator.cpp:189) through the CUDA path: every phonon leaves its wall
    along the normal at 1000 m/s and never scatters.  In the 1000 nm bar each of the 20 sensors (50 nm = 5 measurement
    steps of flight) therefore holds, at every recorded step, the phonons emitted during 5 steps by BOTH walls: N * 5 / 1000
    in total, all carrying +1000 m/s of signed x-velocity (hot phonons move right, cold phonons are negative and move
    left), no y-velocity at all, and the walls absorb everything that arrives.  (Same statement as tests/test_emu.py's
    check of the host build of the device functions.)"""
    model = T.load_model(configs.with_settings(configs.linear(num_phonons=200_000).to_dict(), phasor_sim=True))
    model.prepare()
    src, n = model.sources(4)
    total = sum(src[i].count for i in range(n))
    per_sensor = total * 5 / 1000.0
    for kernel in (2, 0, 1):
        g = psim.GpuSimulator(model.describe(), 0)
        try:
            g.set_option("kernel", kernel)
            g.set_sources(src, n, 4, 0, 1)
            g.run_steps(0, 496)
            alive = g.alive()
            g.run_steps(496, model.info.measurement_steps - 1)
            e, f = g.tallies()
        finally:
            g.close()
        counts = f[:, :, 0] / 1000.0
        assert np.abs(counts - per_sensor).max() <= 3.0, kernel     # stratified births: the count is exact up to rounding
        assert np.abs(f[:, :, 1]).max() == 0.0, kernel
        assert np.abs(e).max() <= 0.05 * per_sensor + 3, kernel     # hot (+1) and cold (-1) phonons cancel in the energy
        assert abs(alive - total / 10) <= 0.01 * total, kernel      # flight time across the bar is 1 ns = 100 steps


@pytest.mark.parametrize("name", ["linear_demo", "sides_trans", "sige"])
def test_phonons_per_cell_are_shard_invariant(name):
    """north star: "the number of phonons per cell" is bit-exact across 1/2/4/8 GPUs.  The histogram over cells of the
    live pool, summed over the shards, after 300 and after 600 measurement steps (windows cut at the same steps, so that
    positions are rounded alike)."""
    model = T.load_model(T.case_model(name), num_phonons=60_000)
    model.prepare()
    src, n = model.sources(7)

    def histograms(shards):
        out = []
        for shard in range(shards):
            g = psim.GpuSimulator(model.describe(), 0)
            try:
                g.set_sources(src, n, 7, shard, shards)
                g.run_steps(0, 300)
                h40 = g.cell_histogram().astype(np.int64)
                a40 = g.alive()
                g.run_steps(300, 600)
                h400 = g.cell_histogram().astype(np.int64)
                assert int(h40.sum()) == a40 and int(h400.sum()) == g.alive()
            finally:
                g.close()
            out.append((h40, h400))
        return sum(h[0] for h in out), sum(h[1] for h in out)

    ref40, ref400 = histograms(1)
    assert ref40.sum() > 0 and ref400.sum() > 0
    for shards in (2, 4, 8):
        h40, h400 = histograms(shards)
        assert np.array_equal(h40, ref40), shards
        assert np.array_equal(h400, ref400), shards


def hot_box(dt_ns: float, num_phonons: int, steps: int = 100):
    """A closed 200 x 200 nm box of Si (specular walls, no emitting surface: nothing ever leaves) whose cells start 5 K above
    the equilibrium temperature, with measurement intervals of dt_ns.  A phonon scatters ~160 times per ns, so with 16 ns
    per interval it consumes ~2600 Philox blocks inside ONE interval (at most ~7000: the fastest-relaxing LA bin at 410 / ns),
    interval after interval; with 100 ns per interval ~16000."""
    m = configs.ModelFile(num_measurements=steps, sim_time=dt_ns * steps, num_phonons=num_phonons, t_eq=300)
    name = m.material(configs.SILICON)
    for i in range(4):
        sid = m.sensor(name, 305.0)
        m.rectangle((i * 50.0, 0.0), ((i + 1) * 50.0, 200.0), sid, 1)
    return m.to_dict()


def test_random_blocks_of_a_long_interval_do_not_repeat():
    """the packed kernels kept 10 bits of Philox block counter per (phonon, measurement step) and saturated
    silently.  Now 13 bits: with 16 ns intervals a phonon consumes well over 1023 blocks per interval (checked: more than
    1500 events per drift-step on average, nearly all of them scatters - the box is one lattice cell until the recorded
    steps begin), and the work-queue and lane-bound kernels (packed counter) must still equal the
    lock-step kernel (counter in a register) bit for bit."""
    model = T.load_model(hot_box(16.0, 400))
    # (lattice_recorded 0: tallies in global memory would otherwise fly the lattice image in the recorded windows as well,
    # staged ones never do - the runs compared here must fly the same cells)
    ref = gpu_run_case(model, 3, steps_per_launch=4, options={"kernel": 1, "tally_shared": 0, "lattice_recorded": 0}, finish=False)
    assert ref["stats"][0]["events"] > 1500 * ref["stats"][0]["drift_steps"]  # or the test is void
    for opts in ({"kernel": 2, "tally_shared": 0}, {"kernel": 0, "tally_shared": 0}, {"kernel": 2, "tally_shared": 1}):
        got = gpu_run_case(model, 3, steps_per_launch=4, options=dict(opts, lattice_recorded=0), finish=False)
        assert np.array_equal(got["energy"], ref["energy"]), opts
        assert np.array_equal(got["fixed"], ref["fixed"]), opts
        assert got["stats"][0]["events"] == ref["stats"][0]["events"], opts


def test_exhausted_random_block_budget_is_an_error_not_a_repeat():
    """Beyond 8191 blocks per (phonon, interval) - 100 ns intervals in the same box - the packed kernels stop with PSIM_E_RNG;
    the lock-step kernel runs the same model to completion."""
    model = T.load_model(hot_box(100.0, 200, steps=20))
    with pytest.raises(psim.PsimError) as err:
        gpu_run_case(model, 3, steps_per_launch=4, options={"kernel": 2}, finish=False)
    assert err.value.code == -8 and "random-number blocks" in err.value.message
    ok = gpu_run_case(model, 3, steps_per_launch=4, options={"kernel": 1}, finish=False)
    assert ok["stats"][0]["events"] > 8191 * ok["stats"][0]["drift_steps"]


def test_many_sensor_tally_path_equals_the_staged_and_the_lane_by_lane_one():
    """linear_sides (1000 sensors): its tallies go to global memory in difference form, posted warp-cooperatively so that
    the three words of an entry share one L2 sector (kernels.cuh:tally_post_global).  The same integers must come out of
    the lane-bound kernel (three REDs per lane) and the lock-step kernel, in periodic and in steady-state mode, for both
    slot counts."""
    for sim_type in (1, 0):
        model = T.load_model(configs.linear_sides(num_phonons=60_000, sim_type=sim_type, step_interval=4 if sim_type else 0).to_dict())
        ref = gpu_run_case(model, 5, steps_per_launch=16, options={"kernel": 1}, finish=False)
        assert ref["stats"][0]["tally_in_shared"] == 3 and np.abs(ref["energy"]).sum() > 0
        for opts in ({"kernel": 2}, {"kernel": 2, "queue_slots": 64}, {"kernel": 2, "queue_slots": 128}, {"kernel": 0}):
            got = gpu_run_case(model, 5, steps_per_launch=16, options=opts, finish=False)
            assert got["stats"][0]["tally_in_shared"] == 3, opts
            assert np.array_equal(got["energy"], ref["energy"]), (sim_type, opts)
            assert np.array_equal(got["fixed"], ref["fixed"]), (sim_type, opts)


def test_rate_class_records_equal_per_sensor_records():
    """The per-phonon loop reads relaxation-rate records by rate CLASS (device_types.h: DevParams::classes).  A model with
    more than 255 distinct sensor temperatures has no classes (255 = unclassified: the sensor's own record); the first 250
    sensors of such a model, alone, have classes.  Both must reproduce the probe of the reference's rate formulas."""
    m = configs.ModelFile(num_measurements=200, sim_time=2, num_phonons=20_000, t_eq=300)
    name = m.material(configs.SILICON)
    n = 300
    for i in range(n):
        sid = m.sensor(name, 300.0 + 0.1 * (i % 290))
        m.rectangle((i * 10.0, 0.0), ((i + 1) * 10.0, 50.0), sid, 1)
    m.emit_surface((0.0, 0.0), (0.0, 50.0), 310)
    m.emit_surface((n * 10.0, 0.0), (n * 10.0, 50.0), 290)
    model = T.load_model(m.to_dict())
    a = gpu_run_case(model, 2, finish=False)
    b = gpu_run_case(model, 2, options={"kernel": 1}, finish=False)
    assert a["stats"][0]["drift_steps"] > 0
    assert np.array_equal(a["energy"], b["energy"]) and np.array_equal(a["fixed"], b["fixed"])


@pytest.mark.parametrize("name", ["linear_demo", "sides_per", "sides_trans"])
def test_reiterated_runs_on_the_device_match_the_reference_with_its_iteration_cap_raised(name):
    """SURVEY.md 8 f2 on the CUDA path: psim_model_run with max_iters = 3 - after every iteration the host takes the new t_eq,
    the sensors' steady temperatures, tables and heat capacities (per measurement step for a transient run: per-(sensor, step)
    records in the device image, device_core.cuh:load_rates) and the GPU simulates again - against fixtures from the
    reference compiled with MAX_ITERS raised to 3 (oracle/Makefile: ref_iters3).  tests/test_emu.py holds the same check for
    the CPU build of the device functions."""
    import os
    from tests import cases
    gold_path = os.path.join(T.GOLDEN, name + ".iters3.npz")
    if not os.path.exists(gold_path):
        pytest.skip("fixture missing")
    gold = np.load(gold_path)
    model = cases.iteration_cases()[name]
    runs = []
    for seed in range(1, 9):
        m = T.load_model(model)
        m.set_max_iters(3)
        st = m.run(device=0, seed=100 * seed)
        six, temps, fluxes = m.results(0)
        info = m.info
        e = np.zeros((info.num_sensors, info.recorded_steps), dtype=np.int32)  # (tallies are not part of this comparison)
        runs.append({"out6": six, "temp_blk": T.blocks(temps, T.NBLOCKS), "flux_blk": T.blocks(fluxes, T.NBLOCKS)})
        assert st.drift_steps > 0 and e.shape[0] > 0
        m.close()
    if model["settings"]["sim_type"] == 0:
        z = T.welch_z(runs, gold, "out6")
        T.assert_parity(z[:, 0], f"{name} x3 temperature column")
        T.assert_parity(z[:, 2], f"{name} x3 x-flux column")
    else:
        T.assert_parity(T.welch_z(runs, gold, "temp_blk"), f"{name} x3 temperature trace")
        T.assert_parity(T.welch_z(runs, gold, "flux_blk"), f"{name} x3 flux trace")


def test_one_iteration_is_the_default_and_three_change_the_answer():
    """max_iters = 1 (the reference as shipped) must stay what every other test checks; with three iterations the steady-state
    bar re-centres t_eq on the mean sensor temperature (309 K between walls at 340 K and 280 K) and its cells emit."""
    from tests import cases
    model = cases.iteration_cases()["linear_demo"]
    one = T.load_model(model)
    s1 = one.run(device=0, seed=5)
    three = T.load_model(model)
    three.set_max_iters(3)
    s3 = three.run(device=0, seed=5)
    a, b = one.results(0)[0], three.results(0)[0]
    assert s3.total_phonons >= s1.total_phonons - 100 and not np.allclose(a[:, 0], b[:, 0], atol=0.05)
    assert abs(a[:, 0].mean() - 300.0) > 5.0  # one iteration: linearised about t_eq = 300 K, the walls are not symmetric about it


def test_triangle_pairs_fly_as_one_parallelogram_without_changing_the_physics():
    """flatten.cpp merges two triangles of one sensor area whose union is a parallelogram into one flight cell (crossing their
    shared edge does nothing to a phonon in the reference either: TransitionSurface::handlePhonon, surface.cpp:71-75).  The
    kinked wire (6174 triangles, 3024 mergeable pairs) must fly through 3150 cells, a mesh cut so that no pair qualifies must keep its triangles,
    the integer bookkeeping that does not depend on trajectories must be identical with and without merging, the per-cell
    histogram must still be in model cells, and 8 seeds each way must agree like two sets of runs of the same code."""
    name = "kinked_spec" if "kinked_spec" in T.all_case_names() else "sides_ss"
    model = T.load_model(T.case_model(name), num_phonons=200_000)
    info = model.info
    merged = [gpu_run_case(model, seed) for seed in range(1, 9)]
    plain = [gpu_run_case(model, seed, options={"merge_cells": 0}) for seed in range(11, 19)]
    assert plain[0]["stats"][0]["flight_cells"] == info.num_cells
    assert merged[0]["stats"][0]["flight_cells"] == (info.num_cells - 3024 if name == "kinked_spec" else info.num_cells // 2)
    same_seed = gpu_run_case(model, 1, options={"merge_cells": 0})
    assert same_seed["sources"] == merged[0]["sources"]
    # far fewer flight segments for the same drift-steps (the shared edges are gone)
    assert merged[0]["stats"][0]["events"] < 0.75 * same_seed["stats"][0]["events"]
    assert abs(merged[0]["stats"][0]["drift_steps"] - same_seed["stats"][0]["drift_steps"]) < 0.01 * same_seed["stats"][0]["drift_steps"]
    feats = {k: [T.run_features(r["energy"], r["flux"], 0, r["six"], r["temps"], r["fluxes"]) for r in runs] for k, runs in (("m", merged), ("p", plain))}
    gold = {"n_seeds": 8}
    for key in ("tally_e", "tally_f", "out6"):
        stack = np.stack([r[key] for r in feats["p"]])
        gold[key + "_mean"], gold[key + "_std"] = stack.mean(axis=0), stack.std(axis=0, ddof=1)
    T.assert_parity(T.welch_z(feats["m"], gold, "tally_e"), "merged vs per-triangle flight cells, energy tallies")
    T.assert_parity(T.welch_z(feats["m"], gold, "tally_f"), "merged vs per-triangle flight cells, flux tallies")
    T.assert_parity(T.welch_z(feats["m"], gold, "out6")[:, 0], "merged vs per-triangle flight cells, temperatures")
    # phonons per MODEL cell after 60 steps: both triangles of the pairs are populated, and the totals agree with the pool
    model.prepare()
    src, n = model.sources(3)
    g = psim.GpuSimulator(model.describe(), 0)
    try:
        g.set_sources(src, n, 3, 0, 1)
        g.run_steps(0, 600 if name == "kinked_spec" else 60)
        hist, alive = g.cell_histogram(), g.alive()
    finally:
        g.close()
    assert int(hist.sum()) == alive and hist.size == info.num_cells
    even, odd = int(hist[0::2].sum()), int(hist[1::2].sum())
    assert abs(even - odd) < 0.1 * alive  # the builder lists the two triangles of a rectangle one after the other


def test_blocks_of_parallelograms_fly_as_one_lattice_cell_where_nothing_is_recorded():
    """flatten.cpp:build_lattices.  In launches that record nothing (the first 90 % of the steps of a steady-state run) every
    rectangular block of identical parallelograms of one material and rate class is ONE flight cell: a transition between
    two of them changes nothing but the sensor label, which such a launch never reads (TransitionSurface::handlePhonon,
    surface.cpp:71-75; the time to scatter need not be redrawn where the rates are the same, modelSimulator.cpp:192-194).
    The pool is converted back to fine flight cells by the first launch that records.  Same seed, with and without: the same
    phonons are emitted, the same random streams drive them, so the runs differ by floating-point rounding only - far less than
    two seeds differ; 8 seeds each way agree like two sets of runs of the same code; the histogram over MODEL cells taken
    while the pool is in lattice coordinates agrees with the one taken without lattices; periodic runs (everything recorded)
    do not use the lattice image at all."""
    from psim_b200 import configs
    for name, steps_mid in (("kinked_spec", 600), ("sige", 300), ("sides_ss", 300), ("linear_rough", 300)):
        if name not in T.all_case_names():
            continue
        model = T.load_model(T.case_model(name), num_phonons=200_000)
        with_l = [gpu_run_case(model, seed) for seed in range(1, 9)]
        without = [gpu_run_case(model, seed, options={"merge_cells": 1}) for seed in range(1, 9)]
        s2, s1 = with_l[0]["stats"][0], without[0]["stats"][0]
        assert s1["lattice_cells"] == 0 and 0 < s2["lattice_cells"] < s2["flight_cells"] == s1["flight_cells"]
        assert with_l[0]["sources"] == without[0]["sources"]
        assert s2["events"] < 0.9 * s1["events"], (name, s2["events"], s1["events"])
        assert abs(s2["drift_steps"] - s1["drift_steps"]) < 0.005 * s1["drift_steps"]
        # same seed: rounding-level differences, a small fraction of the seed-to-seed scatter
        e2 = np.stack([r["energy"].sum(axis=1) for r in with_l]).astype(float)
        e1 = np.stack([r["energy"].sum(axis=1) for r in without]).astype(float)
        assert np.abs(e2 - e1).mean() < 0.35 * e1.std(axis=0, ddof=1).mean(), name
        f2 = np.stack([r["flux"].sum(axis=1) for r in with_l])
        f1 = np.stack([r["flux"].sum(axis=1) for r in without])
        assert np.abs(f2 - f1).mean() < 0.35 * f1.std(axis=0, ddof=1).mean(), name
        # different seeds: two sets of runs of the same physics
        other = [gpu_run_case(model, seed, options={"merge_cells": 1}) for seed in range(11, 19)]
        feats = {k: [T.run_features(r["energy"], r["flux"], 0, r["six"], r["temps"], r["fluxes"]) for r in runs] for k, runs in (("m", with_l), ("p", other))}
        gold = {"n_seeds": 8}
        for key in ("tally_e", "tally_f", "out6"):
            stack = np.stack([r[key] for r in feats["p"]])
            gold[key + "_mean"], gold[key + "_std"] = stack.mean(axis=0), stack.std(axis=0, ddof=1)
        T.assert_parity(T.welch_z(feats["m"], gold, "tally_e"), name + ": lattice vs fine flight cells, energy tallies")
        T.assert_parity(T.welch_z(feats["m"], gold, "tally_f"), name + ": lattice vs fine flight cells, flux tallies")
        # phonons per MODEL cell in the middle of the unrecorded part
        model.prepare()
        hists = {}
        for level in (1, 2):
            src, n = model.sources(3)
            g = psim.GpuSimulator(model.describe(), 0)
            try:
                g.set_option("merge_cells", level)
                g.set_sources(src, n, 3, 0, 1)
                g.run_steps(0, steps_mid)
                hists[level] = (g.cell_histogram(), g.alive())
            finally:
                g.close()
        (h1, a1), (h2, a2) = hists[1], hists[2]
        assert int(h2.sum()) == a2 and int(h1.sum()) == a1 and abs(a1 - a2) <= 0.002 * a1 + 5
        assert np.abs(h2.astype(float) - h1.astype(float)).sum() <= 0.05 * a1 + 20, name  # rounding moves a few phonons across a cell edge
    periodic = T.load_model(configs.linear_sides(sim_type=1, step_interval=4, num_phonons=100_000).to_dict())
    a = gpu_run_case(periodic, 4, finish=False)
    b = gpu_run_case(periodic, 4, options={"merge_cells": 1}, finish=False)
    assert a["stats"][0]["lattice_cells"] > 0  # the image exists ...
    assert np.array_equal(a["energy"], b["energy"]) and np.array_equal(a["fixed"], b["fixed"])  # ... and no launch used it
    assert a["stats"][0]["events"] == b["stats"][0]["events"]


def test_kernel_variants_agree_bit_for_bit_on_a_mesh_that_is_not_axis_aligned():
    """The kinked wire: slanted parallelograms (both terms of every frame product are non-zero), triangles and trapezoids in
    the kinks, composite block edges.  Round 1 / early round 2 left the association of `a b + c d` to the compiler, which
    fused a different product in the lock-step kernel than in the other two: a few flight segments per million went another
    way (device_core.cuh: dot2).  Work queues (2), lane-bound slots (0) and lock step (1) must give identical integers with
    triangles only (merge_cells 0), flight cells (1), lattice cells where nothing is recorded (2) and lattice cells
    throughout (lattice_recorded 1: the work-queue kernel deals the crossed measurements evenly over the lanes,
    kernels.cuh:tally_lattice, the other two walk them lane by lane, device_core.cuh:lattice_runs)."""
    if "kinked_spec" not in T.all_case_names():
        pytest.skip("kinked wire fixture not available")
    model = T.load_model(T.case_model("kinked_spec"), num_phonons=40_000)
    for opts in ({"merge_cells": 0}, {"merge_cells": 1}, {"merge_cells": 2, "lattice_recorded": 0}, {"merge_cells": 2, "lattice_recorded": 1}):
        ref = gpu_run_case(model, 6, options=dict(opts, kernel=1, tally_shared=0), finish=False)
        assert np.abs(ref["energy"]).sum() > 0
        for kern in (2, 0):
            got = gpu_run_case(model, 6, options=dict(opts, kernel=kern, tally_shared=0), finish=False)
            assert np.array_equal(got["energy"], ref["energy"]) and np.array_equal(got["fixed"], ref["fixed"]), (opts, kern)
            assert got["stats"][0]["events"] == ref["stats"][0]["events"] and got["stats"][0]["drift_steps"] == ref["stats"][0]["drift_steps"]


def test_recorded_windows_over_the_lattice_image_attribute_every_measurement_to_its_sensor():
    """Option "lattice_recorded": a flight segment of a lattice cell spans several sensor areas, and the area of every
    measurement it crossed is found from the phonon's position at that instant.  Same seed with and without: the same phonons,
    the same random streams - the per-(sensor, step) tallies differ by rounding only (a phonon within 1e-7 of a cell edge at a
    measurement); far fewer flight segments; the kinked wire uses it by default (its phonons cross 2 - 5 fine cells per step),
    linear_sides does not (half a cell per step)."""
    from psim_b200 import configs
    cases = [("sides_per", T.load_model(configs.linear_sides(sim_type=1, step_interval=4, num_phonons=80_000).to_dict()), False)]
    if "kinked_spec" in T.all_case_names():
        cases.append(("kinked_spec", T.load_model(T.case_model("kinked_spec"), num_phonons=100_000), True))
    for name, model, by_default in cases:
        on = gpu_run_case(model, 8, options={"lattice_recorded": 1}, finish=False)
        off = gpu_run_case(model, 8, options={"lattice_recorded": 0}, finish=False)
        auto = gpu_run_case(model, 8, finish=False)
        assert np.array_equal(auto["energy"], (on if by_default else off)["energy"]), name
        assert on["sources"] == off["sources"]
        assert on["stats"][0]["events"] < 0.85 * off["stats"][0]["events"], name
        e_on, e_off = on["energy"].astype(np.int64), off["energy"].astype(np.int64)
        assert np.abs(e_on - e_off).sum() <= 0.01 * np.abs(e_off).sum(), name
        f_on, f_off = on["fixed"].astype(np.float64), off["fixed"].astype(np.float64)
        assert np.abs(f_on - f_off).sum() <= 0.01 * np.abs(f_off).sum(), name
        assert abs(int(e_on.sum()) - int(e_off.sum())) <= 0.002 * np.abs(e_off).sum() + 10, name
