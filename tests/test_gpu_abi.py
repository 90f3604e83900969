"""Edge cases and error behaviour of the C ABI on a real device (include/psim_b200.h): empty and degenerate inputs,
call-order violations, bad arguments.  The reference signals these with exceptions (std::runtime_error caught in
main.cpp:26); here every entry point returns a negative PSIM_E_* code and psim_gpu_last_error() carries the message."""
import ctypes as C

import numpy as np
import pytest

from psim_b200 import configs, lib as psim
from tests import common as T
from tests.gpu_runner import gpu_run_case

pytestmark = pytest.mark.gpu


def _model(n=20_000, **kw):
    m = T.load_model(configs.linear(num_phonons=n, **kw).to_dict())
    m.prepare()
    return m


def test_no_sources_and_empty_sources_run_to_zero_tallies():
    """A run without phonons is a valid run: zero sources, and sources that all round to zero phonons (emitters at t_eq emit
    nothing but still absorb, modelSimulator.cpp:68-70)."""
    m = _model()
    g = psim.GpuSimulator(m.describe(), 0)
    try:
        src, n = m.sources(1)
        for count in (0, n):  # no source records at all / the model's records with a count of zero
            for i in range(n):
                src[i].count = 0
            g.set_sources(src, count, 1, 0, 1)
            g.run()
            e, f, fx = g.tallies(fixed=True)
            st = g.stats()
            assert not e.any() and not fx.any() and not f.any()
            assert st.total_phonons == 0 and st.drift_steps == 0 and st.events == 0 and st.peak_alive == 0
            assert g.alive() == 0 and not g.cell_histogram().any()
    finally:
        g.close()


def test_ragged_counts_and_single_phonon():
    """Counts that are no multiple of the warp width, down to ONE phonon; every phonon is accounted for: it is either
    still alive at the end or was absorbed by a wall, and its drift-steps are counted."""
    m = _model()
    g = psim.GpuSimulator(m.describe(), 0)
    try:
        src, n = m.sources(1)
        for counts in ((1, 0), (0, 1), (33, 31), (1025, 7)):
            for i in range(n):
                src[i].count = counts[i] if i < len(counts) else 0
            g.set_sources(src, n, 3, 0, 1)
            g.run()
            st = g.stats()
            assert st.total_phonons == sum(counts)
            assert st.shard_phonons == sum(counts)
            assert 0 < st.drift_steps <= sum(counts) * m.info.measurement_steps
            assert g.alive() <= sum(counts)
            e, _ = g.tallies()
            assert np.abs(e).max() <= sum(counts)
    finally:
        g.close()


def test_more_shards_than_phonons():
    """Shards that own no phonon at all (id % num_shards never equals the shard) run and contribute nothing; the shards
    together reproduce the unsharded integers."""
    m = _model()
    src, n = m.sources(1)
    for i in range(n):
        src[i].count = 3 if i == 0 else 0
    g = psim.GpuSimulator(m.describe(), 0)
    try:
        g.set_sources(src, n, 5, 0, 1)
        g.run()
        whole_e, _, whole_f = g.tallies(fixed=True)
        whole_steps = g.stats().drift_steps
        sum_e, sum_f, steps, owners = np.zeros_like(whole_e, dtype=np.int64), np.zeros_like(whole_f), 0, 0
        for shard in range(8):
            g.set_sources(src, n, 5, shard, 8)
            g.run()
            e, _, f = g.tallies(fixed=True)
            st = g.stats()
            owners += int(st.shard_phonons > 0)
            assert st.total_phonons == 3
            sum_e += e
            sum_f += f
            steps += st.drift_steps
        assert owners == 3
        assert np.array_equal(sum_e, whole_e) and np.array_equal(sum_f, whole_f) and steps == whole_steps
    finally:
        g.close()


def test_call_order_and_argument_errors():
    m = _model()
    lib = psim.load_library()
    g = psim.GpuSimulator(m.describe(), 0)
    try:
        with pytest.raises(psim.PsimError) as ei:  # run before set_sources
            g.run()
        assert ei.value.code == -5 and "set_sources" in ei.value.message
        with pytest.raises(psim.PsimError) as ei:
            g.next_window(0)
        assert ei.value.code == -5
        with pytest.raises(psim.PsimError) as ei:
            g.cell_histogram()
        assert ei.value.code == -5
        with pytest.raises(psim.PsimError) as ei:
            g.set_option("no_such_option", 1)
        assert ei.value.code == -1 and "no_such_option" in ei.value.message
        with pytest.raises(psim.PsimError) as ei:
            g.set_option("steps_per_launch", 5000)
        assert ei.value.code == -1
        with pytest.raises(psim.PsimError) as ei:
            g.set_option("queue_slots", 100)
        assert ei.value.code == -5
        src, n = m.sources(1)
        with pytest.raises(psim.PsimError) as ei:  # shard index out of range
            g.set_sources(src, n, 1, 2, 2)
        assert ei.value.code == -1
        bad = (psim.Source * 1)()
        bad[0].kind, bad[0].index, bad[0].sign, bad[0].count = 1, 99, 1, 10  # emitter 99 does not exist
        with pytest.raises(psim.PsimError) as ei:
            g.set_sources(bad, 1, 1, 0, 1)
        assert ei.value.code == -1 and "emitter" in ei.value.message
        bad[0].kind, bad[0].index = 0, 10_000  # nor does cell 10000
        with pytest.raises(psim.PsimError) as ei:
            g.set_sources(bad, 1, 1, 0, 1)
        assert ei.value.code == -1 and "cell" in ei.value.message
        # a failed set_sources leaves the handle without sources, not with half of the old ones
        g.set_sources(src, n, 1, 0, 1)
        with pytest.raises(psim.PsimError):
            g.set_sources(bad, 1, 1, 0, 1)
        with pytest.raises(psim.PsimError) as ei:
            g.run()
        assert ei.value.code == -5
        g.set_sources(src, n, 1, 0, 1)
        with pytest.raises(psim.PsimError) as ei:  # options that shape the pool come before set_sources
            g.set_option("kernel", 0)
        assert ei.value.code == -5
        with pytest.raises(psim.PsimError) as ei:  # steps out of order
            g.run_steps(5, 10)
        assert ei.value.code == -5 and "order" in ei.value.message
        g.run_steps(0, 10)
        with pytest.raises(psim.PsimError):
            g.run_steps(0, 10)  # already done
        g.run()  # continues at step 10
        first = g.tallies(fixed=True)
        g.run()  # nothing left: a second run is a no-op, not an error
        again = g.tallies(fixed=True)
        assert np.array_equal(first[0], again[0]) and np.array_equal(first[2], again[2])
        g.reset()  # back to step 0 with the same sources
        g.run()
        third = g.tallies(fixed=True)
        assert np.array_equal(first[0], third[0]) and np.array_equal(first[2], third[2])
        # NULL handles / pointers
        assert lib.psim_gpu_run(None) == -1 and lib.psim_gpu_reset(None) == -1
        assert lib.psim_gpu_get_stats(g.handle, None) == -1
        assert lib.psim_gpu_next_window(g.handle, 0, None) == -1
        lib.psim_gpu_destroy(None)  # no-op
    finally:
        g.close()
    out = C.c_void_p()
    assert lib.psim_gpu_create(None, 0, C.byref(out)) == -1
    assert lib.psim_gpu_create(m.describe(), 4096, C.byref(out)) == -2 and not out.value
    assert b"device" in lib.psim_gpu_last_error(None)


def test_inconsistent_model_description_is_rejected():
    """psim_gpu_create validates what it copies: indices out of range must not reach the device."""
    m = _model()
    lib = psim.load_library()
    desc = m.describe().contents
    out = C.c_void_p()
    saved = desc.cells[3].sensor
    desc.cells[3].sensor = 10_000
    try:
        assert lib.psim_gpu_create(C.byref(desc), 0, C.byref(out)) == -1 and not out.value
        assert lib.psim_gpu_last_error(None)
    finally:
        desc.cells[3].sensor = saved
    saved = desc.measurement_steps
    desc.measurement_steps = 0
    try:
        assert lib.psim_gpu_create(C.byref(desc), 0, C.byref(out)) == -1 and not out.value
    finally:
        desc.measurement_steps = saved
    assert lib.psim_gpu_create(C.byref(desc), 0, C.byref(out)) == 0
    lib.psim_gpu_destroy(out)


def test_cli_end_to_end_files_and_progress_lines(tmp_path):
    """`psim a.json b.json missing.json` on the device: the reference's progress lines (main.cpp:10-31, model.cpp:141-182),
    `ss_<stem>.txt` with six columns per sensor (outputManager.cpp:72-78) and `per_<stem>.txt` in the three-column block
    format its plotting tools parse (outputManager.cpp:82-114), a failing file reported and skipped, results within the
    reference's golden scatter."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "psim_b200", "bin", "psim")
    a = configs.save(configs.linear(num_phonons=400_000).to_dict(), str(tmp_path / "bar.json"))
    b = configs.save(configs.linear(num_phonons=100_000, sim_type=1, step_interval=4, num_runs=2).to_dict(), str(tmp_path / "wave.json"))
    env = dict(os.environ, PSIM_SEED="11")
    r = subprocess.run([cli, a, str(tmp_path / "missing.json"), b], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0
    out = r.stdout.splitlines()
    assert out[0] == "Run: 1" and out[1].startswith("Stable sensors: ") and out[2] == "System did not stabilize!!"
    assert out[3].startswith("Time Taken: ") and out[3].endswith("[s]")
    assert out[4] == "Run: 1" and "Run: 2" in out and out[-1] == "done"
    assert "missing.json" in r.stderr
    ss = (tmp_path / "ss_bar.txt").read_text().splitlines()
    assert ss[0].startswith('Steady State Results from "bar.json" @ ') and ss[0].endswith("[s] over 1 runs")
    six = np.array([[float(x) for x in line.split()] for line in ss[1:]])
    assert six.shape == (20, 6)
    gold = T.golden("linear_demo")  # 16 reference seeds of this very model: one run must lie within their scatter
    z = (six[:, 0] - gold["out6_mean"][:, 0]) / (gold["out6_std"][:, 0] * np.sqrt(1.0 + 1.0 / int(gold["n_seeds"])))
    assert np.abs(z).max() < 6.0 and np.sqrt((z * z).mean()) < 3.0, z
    per = (tmp_path / "per_wave.txt").read_text().splitlines()
    assert per[0].startswith('Periodic Results from "wave.json" @ ') and per[0].endswith("[s] over 2 runs")
    body = per[1:]
    assert len(body) == 250 * 22  # 1000 steps / interval 4 blocks of (step line, sensor count, 20 sensor rows)
    for blk in (0, 1, 249):
        assert int(body[22 * blk]) == 4 * blk + 2 and int(body[22 * blk + 1]) == 20
        rows = np.array([[float(x) for x in line.split()] for line in body[22 * blk + 2:22 * blk + 22]])
        assert rows.shape == (20, 3) and np.all(rows[:, 0] > 285.0) and np.all(rows[:, 0] < 315.0)


def test_phonon_ids_beyond_32_bits():
    """The same check as tests/test_emu.py:test_phonon_ids_beyond_32_bits on the CUDA path: shard 5999 of 6000 of a
    6e9-phonon job (ids above 2^32: id_hi bits of the packed word and of the Philox counter), and that shard is
    bit-identical whether its million phonons run as one shard or as two half-shards of 12000."""
    from tests.test_emu import _one_of_many_shards_temperatures
    shards = 6000
    model = T.load_model(configs.linear(num_phonons=6_000_000_000).to_dict())
    model.prepare()
    src, n = model.sources(3)
    g = psim.GpuSimulator(model.describe(), 0)
    try:
        def shard_run(k, of):
            g.set_sources(src, n, 3, k, of)
            g.run()
            e, f, fx = g.tallies(fixed=True)
            st = g.stats()
            return e.astype(np.int64), fx, st.drift_steps, st.total_phonons
        e, fx, steps, total = shard_run(shards - 1, shards)
        assert total in range(6_000_000_000 - 2, 6_000_000_000 + 3)
        assert 0.9e6 * 60 < steps < 1.1e6 * 70
        # ids == 5999 (mod 6000) are the ids == 5999 or 11999 (mod 12000)
        e1, f1, s1, _ = shard_run(shards - 1, 2 * shards)
        e2, f2, s2, _ = shard_run(2 * shards - 1, 2 * shards)
        assert np.array_equal(e1 + e2, e) and np.array_equal(f1 + f2, fx) and s1 + s2 == steps
        run = {"energy": e, "flux": fx.astype(np.float64) / 256.0, "sources": [(0, 0, 0, total)]}
        z = _one_of_many_shards_temperatures(run, model, shards)
        assert np.abs(z).max() < 5.0 and np.sqrt((z * z).mean()) < 2.5, z
    finally:
        g.close()


@pytest.mark.gpu
def test_cached_device_memory_is_reused_and_can_be_released():
    """psim_gpu_release_cached (include/psim_b200.h): the device memory of a destroyed handle serves the next handle of the
    process; releasing it in between must change nothing but where the memory comes from - the same run, bit for bit."""
    model = T.load_model(T.case_model("sides_per"), num_phonons=40_000)
    a = gpu_run_case(model, 21, finish=False)
    b = gpu_run_case(model, 21, finish=False)          # pool, tallies, image: all from the cache
    psim.load_library().psim_gpu_release_cached()
    c = gpu_run_case(model, 21, finish=False)          # everything from the driver again
    psim.load_library().psim_gpu_release_cached()
    psim.load_library().psim_gpu_release_cached()      # releasing an empty cache is a no-op
    for other in (b, c):
        assert np.array_equal(a["energy"], other["energy"]) and np.array_equal(a["fixed"], other["fixed"])
        assert a["stats"][0]["drift_steps"] == other["stats"][0]["drift_steps"]
    big = T.load_model(T.case_model("sige"), num_phonons=400_000)   # a larger pool after smaller ones, then a smaller one again
    d = gpu_run_case(big, 5, finish=False)
    e = gpu_run_case(model, 21, finish=False)
    assert np.array_equal(a["energy"], e["energy"]) and np.abs(d["energy"]).sum() > 0
