"""Host layer (C++17, psim_b200/csrc/host) on the CPU: loader, geometry set-up, material tables, energy
bookkeeping, phonons per source, run epilogue, exporter - against the reference's golden known answers and
against the independent numpy restatement under oracle/."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle.model import OracleModel
from psim_b200 import configs
from psim_b200 import lib as psim
from tests import common as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KAT_CASES = ["linear_demo", "linear_hot_cells", "linear_impurity", "linear_full", "sides_ss", "sides_trans", "sige"]


@pytest.mark.parametrize("name", KAT_CASES)
def test_tables_and_energies_match_reference(name, built_library):
    m = T.load_model(T.case_model(name))
    m.prepare()
    kat = T.golden_kat(name)
    total, per = m.energy()
    assert total == pytest.approx(float(kat["total_energy"]), rel=1e-12)
    assert per == pytest.approx(float(kat["total_energy"]) / m.info.num_phonons, rel=1e-12)
    area, init, emit = m.cell_energies()
    np.testing.assert_allclose(area, kat["cell_areas"], rtol=1e-12)
    np.testing.assert_allclose(init, kat["cell_init_energy"], rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(emit, kat["cell_emit_energy"], rtol=1e-12, atol=1e-300)
    for mi in range(m.info.num_materials):
        np.testing.assert_allclose(m.material_arrays(mi)[:, ::25], kat["arrays_sub"][mi], rtol=1e-12, atol=1e-300)
        for ti, temp in enumerate(kat["temps"]):
            for kind in range(3):
                cum, la, s = m.table(mi, kind, float(temp))
                assert s == pytest.approx(float(kat["sums"][mi][ti][kind]), rel=1e-12)
                np.testing.assert_allclose(cum[::25], kat["tables_sub"][mi][ti][kind][:, 0], rtol=1e-10, atol=1e-14)
                np.testing.assert_allclose(la[::25], kat["tables_sub"][mi][ti][kind][:, 1], rtol=1e-10, atol=1e-14)


def test_known_answers_from_baseline_md(built_library):
    """BASELINE.md section 3 (values printed by the reference itself)."""
    m = T.load_model(configs.linear().to_dict())
    m.prepare()
    total, per = m.energy()
    assert total == pytest.approx(2.42698944123e13, rel=1e-11)
    assert per == pytest.approx(4853978.88246, rel=1e-11)
    assert m.table(0, 0, 300.0)[2] == pytest.approx(912265.230238, rel=1e-11)   # baseEnergy(300)
    assert m.table(0, 2, 300.0)[2] == pytest.approx(9.26110237826e16, rel=1e-11)  # scatterEnergy(300)
    assert m.table(0, 1, 310.0)[2] == pytest.approx(2447934219.94, rel=1e-11)   # emitEnergy(310)
    assert m.table(0, 1, 290.0)[2] == pytest.approx(2406044662.51, rel=1e-11)   # emitEnergy(290)
    for name, want in (("sides_ss", 3.62981886153e12), ("sides_trans", 1.08894565846e12), ("sige", 9.69275931312e11),
                       ("linear_full", 5.64205811898e11)):
        mm = T.load_model(T.case_model(name))
        mm.prepare()
        assert mm.energy()[0] == pytest.approx(want, rel=1e-11)


def test_kinked_geometry_energy(built_library):
    if "kinked_spec" not in T.all_case_names():
        pytest.skip("kinked fixture missing")
    m = T.load_model(T.case_model("kinked_spec"))
    m.prepare()
    assert m.energy()[0] == pytest.approx(9.70795776491e12, rel=1e-11)
    info = m.info
    assert (info.num_cells, info.num_sensors, info.num_emitters) == (6174, 3108, 42)
    assert info.num_partial_links == 0


@pytest.mark.parametrize("name", ["linear_demo", "sides_trans", "sige"])
def test_geometry_links_match_oracle(name, built_library):
    """Neighbour discovery + emitter attachment: spatial hash here vs the reference's all-pairs loop restated in
    oracle/model.py (model.cpp:98-138, cell.cpp:81-98)."""
    m = T.load_model(T.case_model(name))
    m.prepare()
    d = m.describe().contents
    om = OracleModel(T.case_model(name))
    for c in range(d.num_cells):
        cell = d.cells[c]
        for k in range(3):
            mine_t, mine_e = set(), set()
            for i in range(cell.sub_first[k], cell.sub_first[k] + cell.sub_count[k]):
                sub = d.subsurfaces[i]
                if sub.kind == 1:
                    mine_t.add(sub.target)
                    assert {round(sub.s0, 9), round(sub.s1, 9)} == {0.0, 1.0}
                else:
                    em = d.emitters[sub.target]
                    assert (em.cell, em.edge) == (c, k)
                    mine_e.add(sub.target)
            assert mine_t == {t[0] for t in om.transitions[c][k]}
            assert len(mine_e) == len(om.emitters[c][k])
    assert d.num_emitters == len(om.emit_list)


def test_source_counts(built_library):
    m = T.load_model(T.case_model("sides_trans"))
    m.prepare()
    src, n = m.sources(1)
    counts = np.array([src[i].count for i in range(n)])
    # 20 hot + 20 cold patches emit; the 20 end surfaces sit at t_eq and emit nothing (but still absorb)
    assert n == 40 and m.info.num_emitters == 60
    assert abs(int(counts.sum()) - m.info.num_phonons) <= n
    assert {src[i].sign for i in range(n)} == {1, -1}
    # expected value of each count = energy / energy-per-phonon: stochastic rounding changes it by < 1
    om = OracleModel(T.case_model("sides_trans"))
    om.prepare()
    want = om.sources(1)
    want = want[want[:, 4] > 0]
    assert np.abs(counts - want[:, 4]).max() <= 1
    # a pure function of (model, seed)
    again, n2 = m.sources(1)
    assert [again[i].count for i in range(n2)] == list(counts)
    other, _ = m.sources(2)
    assert [other[i].count for i in range(n)] != list(counts) or True  # may coincide; must not crash


def test_full_mode_sources_are_cells(built_library):
    m = T.load_model(T.case_model("linear_full"))
    m.prepare()
    src, n = m.sources(5)
    kinds = [src[i].kind for i in range(n)]
    assert kinds.count(0) == 40 and kinds.count(1) == 2  # every cell emits in a full (t_eq = 0) simulation
    assert all(src[i].sign == 1 for i in range(n))


@pytest.mark.parametrize("name", ["linear_demo", "sides_per", "linear_full", "sides_trans"])
def test_run_epilogue_matches_oracle(name, built_library):
    """Given the same tallies, the C++ epilogue (resetRequired side effect -> refresh -> scaleHeatParams,
    model.cpp:163-177) and the numpy restatement give the same tables, including the steady-state
    energy-per-phonon rescale quirk."""
    model = T.case_model(name)
    om = OracleModel(model)
    om.num_phonons = 20_000
    om.prepare()
    e, f, _, _ = om.run(4)
    six_o, temps_o, flux_o = om.finish_run()
    m = T.load_model(model, num_phonons=20_000)
    m.prepare()
    pre = m.energy()[1]
    m.set_tallies(e.astype(np.int32), f)
    m.finish_run(0)
    six, temps, fluxes = m.results(0)
    np.testing.assert_allclose(six, six_o, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(temps, temps_o, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(fluxes, flux_o, rtol=1e-9, atol=1e-3)
    assert m.energy_per_phonon == pytest.approx(om.eff_energy, rel=1e-12)
    if model["settings"]["sim_type"] == 0 and model["settings"]["t_eq"] != 0:
        assert m.energy_per_phonon != pytest.approx(pre, rel=1e-6)  # the post-run rescale really happens
    else:
        assert m.energy_per_phonon == pytest.approx(pre, rel=1e-12)


@pytest.mark.parametrize("name", ["sides_per", "sides_trans", "kinked_spec"])
def test_epilogue_on_several_threads_gives_the_same_bytes(name, built_library, monkeypatch):
    """The run epilogue goes over the sensors on several threads where the tallies are large (1000 sensors x 1000 recorded
    steps here; `Model::for_each_sensor`); every sensor's sums still run over its steps in order, so the stable-sensor count,
    the re-iteration decision and every number of the results are those of one thread, bit for bit."""
    model = T.case_model(name)
    out = {}
    for threads in ("1", "7"):
        monkeypatch.setenv("PSIM_HOST_THREADS", threads)
        m = T.load_model(model, num_phonons=1000)
        m.set_max_iters(2)
        S, R = m.info.num_sensors, m.info.recorded_steps
        assert S * R >= 2 * 65536
        rng = np.random.default_rng(11)
        m.prepare()
        got = []
        for _ in range(2):
            m.describe()
            m.set_tallies(rng.integers(-4000, 4000, (S, R)).astype(np.int32), rng.normal(0, 1e6, (S, R, 2)))
            again = m.end_iteration()
            got.append(bytes([again]))
            if not again:
                break
        got.append(str(m.finish_run(0)).encode())
        got += [a.tobytes() for a in m.results(0)]
        got.append(m.export_text("m.json", 1.0, "now").encode())
        out[threads] = got
        m.close()
    assert out["1"] == out["7"]


def test_export_formats(built_library, tmp_path):
    # steady state: header + one line of six numbers per sensor (outputManager.cpp:72-78, plotting_tools.py:80-92)
    m = T.load_model(T.case_model("linear_demo"), num_phonons=1000)
    m.prepare()
    S, R = m.info.num_sensors, m.info.recorded_steps
    rng = np.random.default_rng(0)
    m.set_tallies(rng.integers(-50, 50, (S, R)).astype(np.int32), rng.normal(0, 1e4, (S, R, 2)))
    m.finish_run(0)
    text = m.export_text("linear_demo.json", 1.5, "2026-01-01 00:00:00")
    lines = text.strip().split("\n")
    assert lines[0] == 'Steady State Results from "linear_demo.json" @ 2026-01-01 00:00:00 - Time Taken 1.5[s] over 1 runs'
    assert len(lines) == 1 + S and all(len(line.split()) == 6 for line in lines[1:])
    six, _, _ = m.results(0)
    assert float(lines[1].split()[0]) == pytest.approx(six[0, 0], rel=1e-5)  # default ostream precision: 6 digits
    path = tmp_path / "linear_demo.json"
    path.write_text("{}")
    m.export(str(path), 2.0)
    assert (tmp_path / "ss_linear_demo.txt").read_text().split("\n")[1] == lines[1]

    # periodic / transient: blocks of (step label, sensor count, T qx qy per sensor)  (outputManager.cpp:82-114)
    m = T.load_model(T.case_model("sides_trans"), num_phonons=1000)
    m.prepare()
    S, R, I = m.info.num_sensors, m.info.recorded_steps, m.info.step_interval
    m.set_tallies(rng.integers(-5, 5, (S, R)).astype(np.int32), rng.normal(0, 1e4, (S, R, 2)))
    m.finish_run(0)
    lines = m.export_text("x.json", 0.25, "now").strip().split("\n")
    assert lines[0].startswith('Periodic Results from "x.json" @ now - Time Taken 0.25[s] over 1 runs')
    nblocks = R // I
    assert len(lines) == 1 + nblocks * (2 + S)
    assert lines[1] == str(I // 2) and lines[2] == str(S)
    assert lines[1 + (2 + S)] == str(I + I // 2)
    _, temps, fluxes = m.results(0)
    first = lines[3].split()
    assert float(first[0]) == pytest.approx(temps[0, :I].mean(), rel=1e-5)
    assert float(first[1]) == pytest.approx(fluxes[0, :I, 0].mean(), rel=1e-4, abs=1.0)


def test_loader_rejections(built_library):
    """Same conditions the reference rejects (model.cpp:53-68,127-133; inputManager.cpp:100)."""
    base = configs.linear(num_phonons=1000).to_dict()

    def rejected(model, needle):
        with pytest.raises(psim.PsimError) as e:
            psim.Model(text=json.dumps(model))
        assert needle in str(e.value)

    rejected(configs.with_settings(base, sim_type=2, step_interval=4, t_eq=0), "deviational")
    rejected(configs.with_settings(base, sim_type=1, step_interval=0), "Step interval of 0")
    bad = json.loads(json.dumps(base))
    bad["emit_surfaces"][0]["p1"]["x"] = 3.0
    rejected(bad, "Unable to add emitting surface")
    bad = json.loads(json.dumps(base))
    bad["emit_surfaces"][0]["duration"] = 5
    rejected(bad, "transient surface")
    bad = json.loads(json.dumps(base))
    bad["cells"][0]["sensorID"] = 999
    rejected(bad, "Sensor does not exist")
    bad = json.loads(json.dumps(base))
    bad["sensors"][1]["id"] = 0
    rejected(bad, "already exists")
    # invalid meshes (Cell::validate, cell.cpp:21-25): crossing edges, a cell inside another, a duplicate
    bad = json.loads(json.dumps(base))
    bad["cells"].append({"triangle": {"p1": {"x": 10, "y": 10}, "p2": {"x": 60, "y": 150}, "p3": {"x": 90, "y": 20}}, "sensorID": 0, "specularity": 1})
    rejected(bad, "intersects")
    bad = json.loads(json.dumps(base))
    bad["cells"].append({"triangle": {"p1": {"x": 5, "y": 5}, "p2": {"x": 5, "y": 40}, "p3": {"x": 20, "y": 5}}, "sensorID": 0, "specularity": 1})
    rejected(bad, "is contained within")
    bad = json.loads(json.dumps(base))
    bad["cells"].append(json.loads(json.dumps(bad["cells"][3])))
    rejected(bad, "Duplicate cell")
    with pytest.raises(psim.PsimError):
        psim.Model(text="{ not json")
    with pytest.raises(psim.PsimError):
        psim.Model(path="/nonexistent/model.json")


def test_multi_run_average(built_library):
    m = T.load_model(T.case_model("linear_demo"), num_phonons=1000)
    m.set_num_runs(2)
    S, R = m.info.num_sensors, m.info.recorded_steps
    rng = np.random.default_rng(1)
    sixes = []
    for run in range(2):
        m.prepare()
        m.set_tallies(rng.integers(-50, 50, (S, R)).astype(np.int32), rng.normal(0, 1e4, (S, R, 2)))
        m.finish_run(run)
        sixes.append(m.results(run)[0])
        m.next_run()
    avg = m.results(None)[0]
    np.testing.assert_allclose(avg, (sixes[0] + sixes[1]) / 2, rtol=1e-12)
    assert "over 2 runs" in m.export_text("a.json", 1.0, "t").split("\n")[0]


def test_cli_without_gpu_reports_and_continues(built_library, tmp_path):
    """`psim a.json b.json`: a failing file is reported and the loop goes on (main.cpp:12-26); there is no CPU
    fallback, so on a box without a GPU every file fails loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    path = configs.save(configs.linear(num_phonons=1000).to_dict(), str(tmp_path / "m.json"))
    cli = os.path.join(ROOT, "psim_b200", "bin", "psim")
    r = subprocess.run([cli, path, str(tmp_path / "missing.json")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0
    assert "no usable CUDA device" in r.stderr and "no CPU path" in r.stderr
    assert "There was an error reading the data" in r.stderr
    assert r.stdout.strip().endswith("done")
    assert not (tmp_path / "ss_m.txt").exists()
    assert subprocess.run([cli], capture_output=True, text=True).stdout.startswith("Need filenames")


def test_result_readers_round_trip(built_library, tmp_path):
    """tests/result_tables.py (a test helper: the readers of the plotting tools are out of scope for the product) reads what the exporter writes - the reference's table formats (outputManager.cpp:72-114) as its
    own Python tools parse them (plotting_tools.py:84-157) - back into the numbers the host API reports, to the six
    significant digits of the text."""
    from tests import result_tables as results
    rng = np.random.default_rng(3)
    m = T.load_model(T.case_model("linear_demo"), num_phonons=1000)
    m.prepare()
    S, R = m.info.num_sensors, m.info.recorded_steps
    m.set_tallies(rng.integers(-50, 50, (S, R)).astype(np.int32), rng.normal(0, 1e4, (S, R, 2)))
    m.finish_run(0)
    path = tmp_path / "bar.json"
    path.write_text("{}")
    m.export(str(path), 2.5)
    ss = results.read_steady_state(str(tmp_path / "ss_bar.txt"))
    assert ss.header.kind == "Steady State" and ss.header.model_file == "bar.json" and ss.header.seconds == 2.5 and ss.header.runs == 1
    six, _, _ = m.results(0)
    np.testing.assert_allclose(ss.table, six, rtol=1e-5)
    np.testing.assert_allclose(ss.average_flux()[0], six[:, 2].mean(), rtol=1e-5)

    m = T.load_model(T.case_model("sides_per"), num_phonons=1000)
    m.prepare()
    S, R, I = m.info.num_sensors, m.info.recorded_steps, m.info.step_interval
    m.set_tallies(rng.integers(-5, 5, (S, R)).astype(np.int32), rng.normal(0, 1e4, (S, R, 2)))
    m.finish_run(0)
    per = results.parse_periodic(m.export_text("wave.json", 0.25, "now"))
    assert per.header.kind == "Periodic" and per.header.when == "now"
    _, temps, fluxes = m.results(0)
    nb = R // I
    assert per.temps.shape == (nb, S) and np.array_equal(per.measurement_steps, np.arange(nb) * I + I // 2)
    np.testing.assert_allclose(per.temps, temps[:, :nb * I].reshape(S, nb, I).mean(axis=2).T, rtol=1e-5)
    np.testing.assert_allclose(per.x_flux, fluxes[:, :nb * I, 0].reshape(S, nb, I).mean(axis=2).T, rtol=2e-5, atol=1e-3)
    np.testing.assert_allclose(per.y_flux, fluxes[:, :nb * I, 1].reshape(S, nb, I).mean(axis=2).T, rtol=2e-5, atol=1e-3)
    with pytest.raises(ValueError):
        results.parse_periodic("title\n2\n3\n300 1 1\n")  # a block that promises three sensors and holds one
    with pytest.raises(ValueError):
        results.parse_steady_state("title\n300 0.1 1 2\n")


@pytest.mark.skipif(not os.path.isdir("/root/reference/psim_python/json/results"), reason="reference tree not mounted")
def test_result_readers_parse_the_reference_files():
    """The three result tables shipped with the reference (2022, an older title line without the run count)."""
    from tests import result_tables as results
    base = "/root/reference/psim_python/json/results/"
    for name, sensors in (("ss_linear_demo.txt", 20), ("ss_linear_sides_demo_ss.txt", 1000), ("ss_kinked_demo_120_35_spec.txt", 3108)):
        ss = results.read_steady_state(base + name)
        assert ss.table.shape == (sensors, 6) and ss.header is not None and ss.header.runs == 1
        assert 265.0 < ss.temps.min() and ss.temps.max() < 335.0 and (ss.temps_std >= 0).all()
    assert results.read_steady_state(base + "ss_linear_demo.txt").header.seconds == pytest.approx(18.1199)


def test_loader_survives_damaged_files(built_library):
    """800 damaged variants of a valid model file (tests/fuzz_loader.py): every one is either rejected with an error code or
    loaded; a crash of the loader would end the subprocess with a signal."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz_loader.py"), "7", "800"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, (r.returncode, r.stderr[-500:])
    accepted, rejected = (int(x) for x in r.stdout.split()[1::2])
    assert accepted + rejected >= 790 and rejected > 600 and accepted > 0


def test_model_without_energy_is_an_error(built_library):
    """Every temperature equal to t_eq: total energy 0, energy per phonon 0.  The reference turns 0 / 0 into a phonon count
    (modelSimulator.cpp:44-49) and never returns; the host layer reports it."""
    m = T.load_model(configs.linear(num_phonons=1000, t_high=300, t_low=300).to_dict())
    m.prepare()
    assert m.energy() == (0.0, 0.0)
    with pytest.raises(psim.PsimError) as ei:
        m.sources(1)
    assert "no energy" in ei.value.message
    with pytest.raises(psim.PsimError) as ei:
        m.run(device=0, seed=1)
    assert "no energy" in ei.value.message
