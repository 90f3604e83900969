"""CPU-side checks of the device algorithm: psim_b200/csrc/device_core.cuh compiled for the host (tests/emu).
The GPU tests run the same checks on the CUDA build; these keep the algorithm honest where no GPU exists."""
import os
from concurrent.futures import ProcessPoolExecutor

import numpy as np
import pytest

from tests import common as T


def _one(args):
    name, seed, n = args
    model = T.load_model(T.case_model(name), num_phonons=n)
    model.prepare()
    r = T.emu_run(model, seed)
    model.set_tallies(r["energy"], r["flux"])
    model.finish_run(0)
    six, temps, fluxes = model.results(0)
    return T.run_features(r["energy"], r["flux"], model.info.sim_type, six, temps, fluxes)


def _one_model(args):
    model_dict, seed = args
    model = T.load_model(model_dict)
    model.prepare()
    r = T.emu_run(model, seed, steps_per_pass=16)
    model.set_tallies(r["energy"], r["flux"])
    model.finish_run(0)
    six, temps, fluxes = model.results(0)
    return T.run_features(r["energy"], r["flux"], model.info.sim_type, six, temps, fluxes)


def _as_gold(runs):
    gold = {"n_seeds": len(runs)}
    for key in ("tally_e", "tally_f", "out6"):
        stack = np.stack([r[key] for r in runs])
        gold[key + "_mean"], gold[key + "_std"] = stack.mean(axis=0), stack.std(axis=0, ddof=1)
    return gold


def test_partial_edges_do_not_change_the_physics():
    """A non-conforming mesh (edges that face two cells: partial transition sub-surfaces) against the same bar meshed
    conformingly - see tests/cases.py:split_bar for why this is a self-consistency check and not a reference parity."""
    from psim_b200 import configs
    from tests import cases
    T.emu_lib()
    cut = cases.split_bar(100_000)
    info = T.load_model(cut).info
    assert info.num_cells == 44 and info.num_partial_links == 8
    plain = configs.linear(num_phonons=100_000).to_dict()
    with ProcessPoolExecutor(8) as ex:
        runs_cut = list(ex.map(_one_model, [(cut, s) for s in range(1, 9)]))
        runs_plain = list(ex.map(_one_model, [(plain, s) for s in range(11, 19)]))
    gold = _as_gold(runs_plain)
    T.assert_parity(T.welch_z(runs_cut, gold, "tally_e"), "cut bar vs plain bar, energy tallies")
    T.assert_parity(T.welch_z(runs_cut, gold, "tally_f"), "cut bar vs plain bar, flux tallies")
    T.assert_parity(T.welch_z(runs_cut, gold, "out6")[:, 0], "cut bar vs plain bar, temperatures")


def _features(name, seeds, n=None):
    T.emu_lib()
    with ProcessPoolExecutor(8) as ex:
        return list(ex.map(_one, [(name, s, n) for s in seeds]))


@pytest.mark.parametrize("name", ["linear_demo", "sige"])
def test_emulated_device_algorithm_steady_state_parity(name):
    gold = T.golden(name)
    runs = _features(name, range(1, 9))
    T.assert_parity(T.welch_z(runs, gold, "tally_e"), f"emu {name} energy tallies")
    T.assert_parity(T.welch_z(runs, gold, "tally_f"), f"emu {name} flux tallies")
    six = T.welch_z(runs, gold, "out6")
    for col in (0, 2, 4):
        T.assert_parity(six[:, col], f"emu {name} ss column {col}")


@pytest.mark.slow
@pytest.mark.parametrize("name", ["sides_trans", "sides_per_full", "kinked_diffuse"])
def test_emulated_device_algorithm_trace_parity(name):
    if name not in T.all_case_names():
        pytest.skip("fixture geometry missing")
    gold = T.golden(name)
    runs = _features(name, range(1, 9))
    T.assert_parity(T.welch_z(runs, gold, "tally_e_blk"), f"emu {name} energy trace")
    T.assert_parity(T.welch_z(runs, gold, "temp_blk"), f"emu {name} temperature trace")
    T.assert_parity(T.welch_z(runs, gold, "flux_blk"), f"emu {name} flux trace")


@pytest.mark.parametrize("name", ["linear_demo", "sides_trans", "linear_full"])
def test_integer_bookkeeping_is_shard_invariant(name):
    """Emitted counts, int32 energy tallies, fixed-point flux tallies, drift-step counts and the per-cell
    population at every step are identical whether 1, 2, 4 or 8 shards share the phonons."""
    model = T.load_model(T.case_model(name), num_phonons=12_000)
    model.prepare()
    ref = T.emu_run(model, 7, want_alive=True, want_hist=True)
    for shards in (2, 4, 8):
        parts = [T.emu_run(model, 7, shard=k, num_shards=shards, want_alive=True, want_hist=True) for k in range(shards)]
        assert all(p["sources"] == ref["sources"] for p in parts)
        assert np.array_equal(sum(p["energy"].astype(np.int64) for p in parts), ref["energy"])
        assert np.array_equal(sum(p["fixed"] for p in parts), ref["fixed"])
        assert sum(p["drift_steps"] for p in parts) == ref["drift_steps"]
        assert np.array_equal(sum(p["alive"] for p in parts), ref["alive"])
        assert np.array_equal(sum(p["hist"] for p in parts), ref["hist"])


def test_steps_per_pass_changes_only_rounding():
    model = T.load_model(T.case_model("sides_per"), num_phonons=8_000)
    model.prepare()
    ref = T.emu_run(model, 3, steps_per_pass=1)
    for spp in (2, 16):
        got = T.emu_run(model, 3, steps_per_pass=spp)
        assert got["sources"] == ref["sources"]
        assert abs(got["drift_steps"] - ref["drift_steps"]) <= 0.02 * ref["drift_steps"]
        e_ref, e_got = np.abs(ref["energy"]).sum(), np.abs(got["energy"]).sum()
        assert abs(float(e_got) - float(e_ref)) <= 0.08 * float(e_ref)


def test_birth_bookkeeping():
    """Every phonon of every source is created exactly once (or falls in the unrecorded last interval), windows of
    transient surfaces are honoured, and the pool population never exceeds what was emitted."""
    model = T.load_model(T.case_model("sides_trans"), num_phonons=20_000)
    model.prepare()
    r = T.emu_run(model, 5, want_alive=True)
    total = sum(c for _, _, _, c in r["sources"])
    assert abs(total - 20_000) <= len(r["sources"])
    alive = r["alive"].astype(np.int64)
    assert alive.max() <= total
    M = model.info.measurement_steps
    # transient patches emit in [0.1, 0.25] and [0.25, 0.4] ns of 0.5 ns: nothing is alive before step 200
    assert alive[: int(0.1 / 0.5 * M) - 1].max() == 0
    assert alive[int(0.2 / 0.5 * M)] > 0


def test_full_mode_initial_population():
    model = T.load_model(T.case_model("linear_full"), num_phonons=10_000)
    model.prepare()
    r = T.emu_run(model, 2, want_alive=True, want_hist=True)
    cells = sum(c for k, _, _, c in r["sources"] if k == 0)
    assert cells > 100                    # in a full simulation every cell starts with its thermal population
    assert r["alive"][0] >= 0.9 * cells   # which is alive after the first interval
    assert (r["hist"][0] > 0).sum() >= 38  # in (nearly) every one of the 40 cells


def test_phasor_mode_is_exact_kinematics():
    """phasor_sim (phononBuilder.cpp:42-49, modelSimulator.cpp:189): every phonon leaves its wall along the normal at
    1000 m/s and never scatters.  In the 1000 nm bar each of the 20 sensors (50 nm = 5 measurement steps of flight)
    therefore holds, at every recorded step, the phonons emitted during 5 steps by BOTH walls: N * 5 / 1000 in
    total, all carrying +1000 m/s of signed x-velocity (hot phonons move right, cold phonons are negative and move
    left), no y-velocity at all, and the walls absorb everything that arrives."""
    from psim_b200 import configs
    model = configs.with_settings(configs.linear(num_phonons=20_000).to_dict(), phasor_sim=True)
    m = T.load_model(model)
    m.prepare()
    r = T.emu_run(m, 4, steps_per_pass=16, want_alive=True)
    n = sum(c for _, _, _, c in r["sources"])
    per_sensor = n * 5 / 1000.0
    counts = r["flux"][:, :, 0] / 1000.0
    assert np.abs(counts - per_sensor).max() <= 3.0
    assert np.abs(r["flux"][:, :, 1]).max() == 0.0
    assert np.abs(r["energy"]).max() <= 0.05 * per_sensor + 3
    # flight time across the bar is 1 ns = 100 steps: the population saturates at n / 10
    assert abs(int(r["alive"][16 * 31 - 1]) - n / 10) <= 0.01 * n  # population is recorded at the end of each 16-step pass


def _one_of_many_shards_temperatures(run, model, shards):
    """Temperatures from ONE shard's tallies scaled by the number of shards (deviational mode: every shard is an unbiased
    sample of the same job), as z against the reference's fixture of the same bar."""
    model.set_tallies((run["energy"].astype(np.int64) * shards).astype(np.int32), run["flux"] * shards)
    model.finish_run(0)
    six, _, _ = model.results(0)
    gold = T.golden("linear_demo")
    reduced = T.case_model("linear_demo")["settings"]["num_phonons"]
    mine = sum(c for _, _, _, c in run["sources"]) / shards
    se = gold["out6_std"][:, 0] * np.sqrt(reduced / mine + 1.0 / int(gold["n_seeds"]))
    return (six[:, 0] - gold["out6_mean"][:, 0]) / se


def test_phonon_ids_beyond_32_bits():
    """6e9 phonons: global phonon ids need more than 32 bits (id_lo + the 8 id_hi bits of the packed word, which also key
    the Philox counter).  Shard 5999 of 6000 simulates the million ids congruent to 5999, most of them above 2^32; its
    tallies scaled by 6000 must give the temperatures of the bar."""
    from psim_b200 import configs
    shards = 6000
    model = T.load_model(configs.linear(num_phonons=6_000_000_000).to_dict())
    model.prepare()
    run = T.emu_run(model, 3, shard=shards - 1, num_shards=shards, steps_per_pass=16)
    assert sum(c for _, _, _, c in run["sources"]) in range(6_000_000_000 - 2, 6_000_000_000 + 3)
    assert 0.9e6 * 60 < run["drift_steps"] < 1.1e6 * 70  # a million phonons at ~65 drift-steps each
    z = _one_of_many_shards_temperatures(run, model, shards)
    assert np.abs(z).max() < 5.0 and np.sqrt((z * z).mean()) < 2.5, z


def test_damaged_descriptions_are_rejected_or_run_to_completion():
    """1500 damaged psim_model_desc / psim_source inputs (tests/fuzz_desc.py) through flatten_model + plan_births and, when
    accepted, a run of the emulated particle loop: error code or finished run, never a crash (the subprocess would end
    with a signal) or a hang (timeout)."""
    import os
    import subprocess
    import sys
    T.emu_lib()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # a transient model (every pass records: fine flight cells), then a steady-state one (its unrecorded passes fly the
    # lattice image, flatten.cpp:build_lattices, and the first recording pass converts the pool back)
    for env, iterations in (({}, 1500), ({"FUZZ_STEADY": "1"}, 700)):
        r = subprocess.run([sys.executable, os.path.join(root, "tests", "fuzz_desc.py"), "5", str(iterations)], capture_output=True, text=True,
                           timeout=600, env=dict(os.environ, **env))
        assert r.returncode == 0, (env, r.returncode, r.stderr[-500:])
        first, last = r.stdout.strip().splitlines()[0], r.stdout.strip().splitlines()[-1]
        assert first.startswith("baseline (0,") and "final (0," in last  # the undamaged description runs before and after
        errors, ok = int(last.split()[1]), int(last.split()[3])
        assert errors > iterations // 3 and ok > iterations // 5


def _iterated_features(model_dict, seed, engine):
    """One run of a model whose iteration cap is 3, driven the way psim_model_run drives it: simulate, end the iteration,
    and - while the model asks for it - describe the model again and simulate again.  `engine(model, seed)` runs the
    particle loop once and returns (energy, flux)."""
    m = T.load_model(model_dict)
    m.set_max_iters(3)
    m.prepare()
    iterations = 0
    while True:
        e, f = engine(m, seed + 7919 * iterations)
        iterations += 1
        m.set_tallies(e, f)
        if not m.end_iteration():
            break
    m.finish_run(0)
    six, temps, fluxes = m.results(0)
    feats = T.run_features(e, f, m.info.sim_type, six, temps, fluxes)
    m.close()
    return feats, iterations


@pytest.mark.parametrize("name", ["linear_demo", "sides_per", "sides_trans"])
def test_reiterated_runs_match_the_reference_with_its_iteration_cap_raised(name):
    """the re-iteration of a run (model.cpp:159-172: new t_eq, sensor steady temperatures, tables and heat
    capacities from the iteration before; per measurement step for a transient run, sensorController.cpp:101-113).  Upstream
    MAX_ITERS = 1 makes it dead code, so the fixture comes from the reference compiled with that one constant raised to 3
    (oracle/Makefile: ref_iters3; tests/golden/make_golden.py --iters3).  Here the host logic of psim_b200 (end_iteration,
    describe with per-sensor and per-step records) around the CPU build of the device functions."""
    from tests import cases
    gold_path = os.path.join(T.GOLDEN, name + ".iters3.npz")
    if not os.path.exists(gold_path):
        pytest.skip("fixture missing")
    gold = np.load(gold_path)
    model = cases.iteration_cases()[name]

    def engine(m, seed):
        r = T.emu_run(m, seed, steps_per_pass=16)
        return r["energy"], r["flux"]

    runs = []
    for seed in range(1, 7):
        feats, iterations = _iterated_features(model, seed, engine)
        assert iterations == 3
        runs.append(feats)
    if model["settings"]["sim_type"] == 0:
        z = T.welch_z(runs, gold, "out6")
        T.assert_parity(z[:, 0], f"emu {name} x3 temperature column")
        T.assert_parity(z[:, 2], f"emu {name} x3 x-flux column")
    else:
        T.assert_parity(T.welch_z(runs, gold, "temp_blk"), f"emu {name} x3 temperature trace")
        T.assert_parity(T.welch_z(runs, gold, "flux_blk"), f"emu {name} x3 flux trace")


def test_flatten_pairs_triangles_into_parallelograms_and_dedupes_shapes():
    """flatten.cpp: two triangles of one sensor area whose union is a parallelogram become ONE flight cell (device_types.h),
    and cell geometry is stored once per distinct shape.  The kinked wire: 6174 triangles -> 3087 flight cells; every shipped
    mesh halves; a bar whose rectangles belong to two sensors each (no pair inside one sensor area) keeps its triangles."""
    import ctypes as C
    from psim_b200 import configs
    from tests import cases
    lib = T.emu_lib()
    lib.psim_emu_mesh_info.restype = C.c_int

    def info(model_dict, merge=1):
        m = T.load_model(model_dict)
        m.prepare()
        n, s, k = C.c_uint32(), C.c_uint32(), C.c_uint32()
        assert lib.psim_emu_mesh_info(m.describe(), merge, C.byref(n), C.byref(s), C.byref(k)) == 0
        return m.info.num_cells, n.value, s.value, k.value

    for model in (configs.linear(num_phonons=1000).to_dict(), configs.linear_sides(num_phonons=1000).to_dict(),
                  configs.si_ge_grid(num_phonons=1000).to_dict(), cases.split_bar(1000)):
        cells, flight, shapes, classes = info(model)
        assert flight * 2 == cells and shapes <= 2 and classes <= 2
        assert info(model, merge=0)[1] == cells
    kinked = cases.kinked_model()
    if kinked is not None:
        cells, flight, shapes, classes = info(configs.with_settings(kinked, num_phonons=1000))
        # 3024 of its 3087 rectangles are parallelograms inside one sensor area (42 in the kinks are trapezoids, 21 are shared by two sensors)
        assert (cells, flight) == (6174, 6174 - 3024) and shapes <= 160 and classes == 1
    # two sensors per rectangle: the triangles of a rectangle lie in different sensor areas and must not be merged
    m = configs.ModelFile(num_measurements=100, sim_time=1, num_phonons=1000, t_eq=300)
    name = m.material(configs.SILICON)
    for i in range(4):
        a, b = m.sensor(name, 300.0), m.sensor(name, 300.0)
        m.triangle((i * 10.0, 0.0), (i * 10.0, 10.0), ((i + 1) * 10.0, 10.0), a, 1)
        m.triangle((i * 10.0, 0.0), ((i + 1) * 10.0, 10.0), ((i + 1) * 10.0, 0.0), b, 1)
    m.emit_surface((0.0, 0.0), (0.0, 10.0), 310)
    m.emit_surface((40.0, 0.0), (40.0, 10.0), 290)
    cells, flight, _, _ = info(m.to_dict())
    assert cells == 8 and flight == 8


def test_lattice_image_blocks_and_their_neutrality():
    """flatten.cpp:build_lattices - the image flown by the launches that record nothing: every rectangular block of identical
    parallelograms of one material and rate class is one cell.  The Si/Ge bench model is two blocks of 5 x 5 (silicon half,
    germanium half), linear_sides one block of 100 x 10, the bar one of 20 x 1, the kinked wire 250 cells (118 blocks, the
    largest 21 x 21, plus the triangles and trapezoids of its kinks); a mesh whose sensors all differ in temperature has no
    block, and neither has a re-iterated transient run (per-step rates).  Neutrality: with the same seed the same phonons are
    emitted and driven by the same random streams, so a run with the lattice image differs from one without by floating-point
    rounding only - a small fraction of what two seeds differ by - and far fewer flight segments are flown."""
    import ctypes as C
    from psim_b200 import configs
    from tests import cases
    lib = T.emu_lib()
    lib.psim_emu_lattice_info.restype = C.c_int

    def info(model_dict):
        m = T.load_model(model_dict)
        m.prepare()
        v = [C.c_uint32() for _ in range(4)]
        assert lib.psim_emu_lattice_info(m.describe(), *[C.byref(x) for x in v]) == 0
        return tuple(x.value for x in v)  # cells, blocks of more than one, largest block, partial-edge records

    assert info(configs.si_ge_grid(num_phonons=1000).to_dict())[:3] == (2, 2, 25)
    assert info(configs.linear_sides(num_phonons=1000).to_dict())[:3] == (1, 1, 1000)
    assert info(configs.linear(num_phonons=1000).to_dict())[:3] == (1, 1, 20)
    kinked = cases.kinked_model()
    if kinked is not None:
        cells, merged, largest, subs = info(configs.with_settings(kinked, num_phonons=1000))
        assert (cells, merged, largest) == (250, 118, 441) and subs < 4000
    graded = configs.linear(num_phonons=1000).to_dict()
    graded["sensors"] = [dict(s, t_init=300.0 + 0.1 * i) for i, s in enumerate(graded["sensors"])]
    assert info(graded)[:3] == (0, 0, 0)  # every sensor its own rate class: nothing to merge, no image
    try:
        for name, spp in (("sige", 64), ("sides_ss", 64), ("linear_rough", 64)):
            model = T.load_model(T.case_model(name), num_phonons=30_000)
            model.prepare()
            runs = {}
            for level in (1, 2):
                lib.psim_emu_set_merge_cells(level)
                runs[level] = [T.emu_run(model, seed, steps_per_pass=spp) for seed in (1, 2, 3, 4)]
            e1 = np.stack([r["energy"].sum(axis=1) for r in runs[1]]).astype(float)
            e2 = np.stack([r["energy"].sum(axis=1) for r in runs[2]]).astype(float)
            assert np.abs(e2 - e1).mean() < 0.25 * e1.std(axis=0, ddof=1).mean(), name
            for a, b in zip(runs[1], runs[2]):
                assert a["sources"] == b["sources"]
                assert abs(a["drift_steps"] - b["drift_steps"]) < 0.002 * a["drift_steps"], name
                assert b["events"] < 0.9 * a["events"], name
    finally:
        lib.psim_emu_set_merge_cells(2)


def test_recorded_passes_over_the_lattice_image_attribute_every_measurement_to_its_sensor():
    """device_core.cuh:lattice_runs / lattice_sensor_at and flatten.cpp's sub_sensor table, on the CPU: with recorded passes
    over the lattice image a flight segment spans several sensor areas and the area of every measurement it crossed is found
    from the position at that instant.  Same seed with and without: the per-(sensor, step) tallies differ by rounding only
    (a phonon within 1e-7 of a cell edge at a measurement) while far fewer segments are flown; the library's rule switches it
    on for the kinked wire (4.9 fine cells per step at the largest group velocity) and leaves linear_sides alone (0.46)."""
    import ctypes as C
    from psim_b200 import configs
    lib = T.emu_lib()
    cases = [("sides_per", configs.linear_sides(sim_type=1, step_interval=4, num_phonons=20_000).to_dict(), 64, False)]
    if "kinked_spec" in T.all_case_names():
        cases.append(("kinked_spec", configs.with_settings(T.case_model("kinked_spec"), num_phonons=20_000), 128, True))
    try:
        for name, model_dict, spp, by_default in cases:
            model = T.load_model(model_dict)
            model.prepare()
            runs = {}
            for setting in (0, 1, -1):
                lib.psim_emu_set_lattice_recorded(setting)
                runs[setting] = T.emu_run(model, 2, steps_per_pass=spp)
            off, on, auto = runs[0], runs[1], runs[-1]
            assert np.array_equal(auto["energy"], (on if by_default else off)["energy"]), name
            assert on["sources"] == off["sources"] and on["events"] < 0.85 * off["events"], name
            e_on, e_off = on["energy"].astype(np.int64), off["energy"].astype(np.int64)
            assert np.abs(e_on - e_off).sum() <= 0.01 * np.abs(e_off).sum(), name
            assert np.abs(on["fixed"] - off["fixed"]).sum() <= 0.01 * np.abs(off["fixed"]).sum(), name
    finally:
        lib.psim_emu_set_lattice_recorded(-1)
