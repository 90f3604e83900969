"""ctypes bindings of include/psim_b200.h and include/psim_host.h.

``Model`` mirrors the reference's ``Model`` as its ``main`` drives it (reference psim/src/main.cpp:15-21:
deserialize -> runSimulation -> exportResults); ``GpuSimulator`` mirrors ``ModelSimulator``
(psim/include/psim/modelSimulator.h:12-41: initPhononBuilders / runSimulation / reset).  All arithmetic of the
particle loop happens in libpsim_b200.so on the GPU; this module only moves buffers.  If the library cannot be
loaded or no CUDA device is usable the calls raise - there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PSIM_B200_LIB") or os.path.join(_HERE, "lib", "libpsim_b200.so")  # env: tuning builds only
BINS = 1000

ERRORS = {0: "OK", -1: "PSIM_E_INVALID", -2: "PSIM_E_NO_DEVICE", -3: "PSIM_E_CUDA", -4: "PSIM_E_OVERFLOW",
          -5: "PSIM_E_STATE", -6: "PSIM_E_IO", -7: "PSIM_E_MODEL", -8: "PSIM_E_RNG"}


class PsimError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{ERRORS.get(code, code)}: {message}")
        self.code = code
        self.message = message


class Material(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("b_l", "b_tn", "b_tu", "b_i", "w", "w_max_la", "w_max_ta", "freq_width")]


class Sensor(C.Structure):
    _fields_ = [("material", C.c_uint32), ("base_table", C.c_uint32), ("scatter_table", C.c_uint32),
                ("reserved", C.c_uint32), ("temperature", C.c_double)]


class SubSurface(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("target", C.c_uint32), ("target_edge", C.c_uint32), ("reserved", C.c_uint32),
                ("s0", C.c_double), ("s1", C.c_double), ("t0", C.c_double), ("t1", C.c_double)]


class Cell(C.Structure):
    _fields_ = [("x", C.c_double * 3), ("y", C.c_double * 3), ("specularity", C.c_double), ("sensor", C.c_uint32),
                ("sub_first", C.c_uint32 * 3), ("sub_count", C.c_uint32 * 3), ("reserved", C.c_uint32)]


class Emitter(C.Structure):
    _fields_ = [("cell", C.c_uint32), ("edge", C.c_uint32), ("table", C.c_uint32), ("reserved", C.c_uint32),
                ("s_p1", C.c_double), ("s_p2", C.c_double), ("start_time", C.c_double), ("duration", C.c_double)]


class Table(C.Structure):
    _fields_ = [("cumulative", C.POINTER(C.c_double)), ("la_fraction", C.POINTER(C.c_double))]


class ModelDesc(C.Structure):
    _fields_ = [("num_materials", C.c_uint32), ("num_sensors", C.c_uint32), ("num_cells", C.c_uint32),
                ("num_subsurfaces", C.c_uint32), ("num_emitters", C.c_uint32), ("num_tables", C.c_uint32),
                ("materials", C.POINTER(Material)), ("velocities", C.POINTER(C.c_double)),
                ("sensors", C.POINTER(Sensor)), ("cells", C.POINTER(Cell)), ("subsurfaces", C.POINTER(SubSurface)),
                ("emitters", C.POINTER(Emitter)), ("tables", C.POINTER(Table)),
                ("measurement_steps", C.c_uint32), ("step_adjustment", C.c_uint32), ("simulation_time", C.c_double),
                ("full_simulation", C.c_uint32), ("phasor_sim", C.c_uint32), ("step_sensors", C.POINTER(Sensor))]


class Source(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("index", C.c_uint32), ("sign", C.c_int32), ("reserved", C.c_uint32),
                ("count", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("total_phonons", C.c_uint64), ("shard_phonons", C.c_uint64), ("drift_steps", C.c_uint64),
                ("events", C.c_uint64), ("peak_alive", C.c_uint64), ("kernel_ms", C.c_double),
                ("launches", C.c_uint32), ("steps_per_launch", C.c_uint32), ("warps", C.c_uint32),
                ("tally_in_shared", C.c_uint32), ("image_bytes", C.c_uint64), ("plan_bytes", C.c_uint64),
                ("tally_bytes", C.c_uint64), ("kernel", C.c_uint32), ("flight_cells", C.c_uint32),
                ("lattice_cells", C.c_uint32), ("lattice_recorded", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class ModelInfo(C.Structure):
    _fields_ = [("num_runs", C.c_uint64), ("measurement_steps", C.c_uint64), ("recorded_steps", C.c_uint64),
                ("num_phonons", C.c_uint64), ("step_interval", C.c_uint64), ("simulation_time", C.c_double),
                ("t_eq", C.c_double), ("sim_type", C.c_uint32), ("phasor_sim", C.c_uint32),
                ("num_materials", C.c_uint32), ("num_sensors", C.c_uint32), ("num_cells", C.c_uint32),
                ("num_emitters", C.c_uint32), ("num_transition_links", C.c_uint32), ("num_partial_links", C.c_uint32)]


# every symbol include/*.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_DP = C.POINTER(C.c_double)
SYMBOLS = {
    # include/psim_b200.h
    "psim_gpu_create": (C.c_int, [C.POINTER(ModelDesc), C.c_int, C.POINTER(_P)]),
    "psim_gpu_set_sources": (C.c_int, [_P, C.POINTER(Source), C.c_size_t, C.c_uint64, C.c_uint32, C.c_uint32]),
    "psim_gpu_run": (C.c_int, [_P]),
    "psim_gpu_run_steps": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P]),
    "psim_gpu_synchronize": (C.c_int, [_P]),
    "psim_gpu_next_window": (C.c_int, [_P, C.c_uint32, C.POINTER(C.c_uint32)]),
    "psim_gpu_get_tallies": (C.c_int, [_P, _P, _P, _P]),
    "psim_gpu_tally_buffers": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "psim_gpu_alive": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "psim_gpu_cell_histogram": (C.c_int, [_P, _P]),
    "psim_gpu_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "psim_gpu_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "psim_gpu_reset": (C.c_int, [_P]),
    "psim_gpu_destroy": (None, [_P]),
    "psim_gpu_last_error": (C.c_char_p, [_P]),
    "psim_gpu_release_cached": (None, []),
    "psim_gpu_probe_sample": (C.c_int, [_P, C.c_uint32, _P, _P, C.c_size_t, _P, _P, _P]),
    "psim_gpu_probe_rates": (C.c_int, [_P, C.c_uint32, _P, _P, C.c_size_t, _P]),
    "psim_gpu_probe_flight": (C.c_int, [_P, _P, _P, C.c_size_t, _P]),
    # include/psim_host.h
    "psim_host_last_error": (C.c_char_p, []),
    "psim_model_load": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "psim_model_load_text": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "psim_model_free": (None, [_P]),
    "psim_model_get_info": (C.c_int, [_P, C.POINTER(ModelInfo)]),
    "psim_model_set_num_phonons": (C.c_int, [_P, C.c_uint64]),
    "psim_model_set_num_runs": (C.c_int, [_P, C.c_uint64]),
    "psim_model_set_max_iters": (C.c_int, [_P, C.c_uint64]),
    "psim_model_end_iteration": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "psim_model_prepare": (C.c_int, [_P]),
    "psim_model_energy": (C.c_int, [_P, _DP, _DP]),
    "psim_model_material_arrays": (C.c_int, [_P, C.c_uint32, _P, _P, _P, _P, _P]),
    "psim_model_table": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_double, _P, _P, _DP]),
    "psim_model_cell_energies": (C.c_int, [_P, _P, _P, _P]),
    "psim_model_sensor_ids": (C.c_int, [_P, _P, _P]),
    "psim_model_describe": (C.c_int, [_P, C.POINTER(C.POINTER(ModelDesc))]),
    "psim_model_sources": (C.c_int, [_P, C.c_uint64, C.POINTER(Source), C.POINTER(C.c_size_t)]),
    "psim_model_set_tallies": (C.c_int, [_P, _P, _P]),
    "psim_model_finish_run": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_int)]),
    "psim_model_next_run": (C.c_int, [_P]),
    "psim_model_run": (C.c_int, [_P, C.c_int, C.c_uint64, C.c_int, C.c_int, C.POINTER(Stats)]),
    "psim_model_run_devices": (C.c_int, [_P, C.POINTER(C.c_int), C.c_int, C.c_uint64, C.c_int, C.c_int, C.POINTER(Stats)]),
    "psim_model_results": (C.c_int, [_P, C.c_uint64, _P, _P, _P]),
    "psim_model_energy_per_phonon": (C.c_double, [_P]),
    "psim_model_export": (C.c_int, [_P, C.c_char_p, C.c_double]),
    "psim_model_export_text": (C.c_int, [_P, C.c_char_p, C.c_double, C.c_char_p, C.c_char_p, C.POINTER(C.c_size_t)]),
}

_lib = None


def load_library() -> C.CDLL:
    """Load libpsim_b200.so (built in-tree by psim_b200/build.py); raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PsimError(-2, f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Model:
    """Host-side model: loader, tables, energy bookkeeping, run epilogue, exporter (CPU)."""

    def __init__(self, path: Optional[str] = None, text: Optional[str] = None):
        self.lib = load_library()
        self.handle = C.c_void_p()
        if path is not None:
            rc = self.lib.psim_model_load(os.fsencode(path), C.byref(self.handle))
        elif text is not None:
            rc = self.lib.psim_model_load_text(text.encode(), C.byref(self.handle))
        else:
            raise ValueError("path or text required")
        if rc:
            raise PsimError(rc, self.lib.psim_host_last_error().decode())
        self.path = path
        self.runs_done = 0

    def close(self):
        if getattr(self, "handle", None):
            self.lib.psim_model_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc:
            raise PsimError(rc, self.lib.psim_host_last_error().decode())

    @property
    def info(self) -> ModelInfo:
        out = ModelInfo()
        self._check(self.lib.psim_model_get_info(self.handle, C.byref(out)))
        return out

    def set_num_phonons(self, n: int):
        self._check(self.lib.psim_model_set_num_phonons(self.handle, int(n)))

    def set_num_runs(self, n: int):
        self._check(self.lib.psim_model_set_num_runs(self.handle, int(n)))

    def set_max_iters(self, n: int):
        """Iterations per run (the reference's MAX_ITERS, model.cpp:11)."""
        self._check(self.lib.psim_model_set_max_iters(self.handle, int(n)))

    def end_iteration(self) -> bool:
        """model.cpp:163-171; True if the run must be simulated again with a fresh describe() / sources()."""
        again = C.c_int()
        self._check(self.lib.psim_model_end_iteration(self.handle, C.byref(again)))
        return bool(again.value)

    def prepare(self):
        self._check(self.lib.psim_model_prepare(self.handle))

    def energy(self) -> Tuple[float, float]:
        tot, per = C.c_double(), C.c_double()
        self._check(self.lib.psim_model_energy(self.handle, C.byref(tot), C.byref(per)))
        return tot.value, per.value

    def material_arrays(self, material: int) -> np.ndarray:
        out = np.zeros((5, BINS))
        self._check(self.lib.psim_model_material_arrays(self.handle, material, *[_ptr(out[i]) for i in range(5)]))
        return out

    def table(self, material: int, kind: int, temperature: float):
        cum, la, s = np.zeros(BINS), np.zeros(BINS), C.c_double()
        self._check(self.lib.psim_model_table(self.handle, material, kind, float(temperature), _ptr(cum), _ptr(la), C.byref(s)))
        return cum, la, s.value

    def cell_energies(self):
        n = self.info.num_cells
        a, i, e = np.zeros(n), np.zeros(n), np.zeros(n)
        self._check(self.lib.psim_model_cell_energies(self.handle, _ptr(a), _ptr(i), _ptr(e)))
        return a, i, e

    def sensor_ids(self):
        n = self.info.num_sensors
        ids, areas = np.zeros(n, dtype=np.uint64), np.zeros(n)
        self._check(self.lib.psim_model_sensor_ids(self.handle, _ptr(ids), _ptr(areas)))
        return ids, areas

    def describe(self) -> "C.POINTER(ModelDesc)":
        out = C.POINTER(ModelDesc)()
        self._check(self.lib.psim_model_describe(self.handle, C.byref(out)))
        return out

    def sources(self, seed: int):
        n = C.c_size_t(0)
        self._check(self.lib.psim_model_sources(self.handle, int(seed), None, C.byref(n)))
        arr = (Source * max(n.value, 1))()
        n2 = C.c_size_t(n.value)
        self._check(self.lib.psim_model_sources(self.handle, int(seed), arr, C.byref(n2)))
        return arr, n2.value

    def set_tallies(self, energy: np.ndarray, flux: np.ndarray):
        energy = np.ascontiguousarray(energy, dtype=np.int32)
        flux = np.ascontiguousarray(flux, dtype=np.float64)
        self._check(self.lib.psim_model_set_tallies(self.handle, _ptr(energy), _ptr(flux)))

    def finish_run(self, run_id: int = 0) -> int:
        stable = C.c_int()
        self._check(self.lib.psim_model_finish_run(self.handle, run_id, C.byref(stable)))
        self.runs_done = max(self.runs_done, run_id + 1)
        return stable.value

    def next_run(self):
        self._check(self.lib.psim_model_next_run(self.handle))

    def run(self, device: int = 0, seed: int = 1, steps_per_launch: int = 0, verbose: bool = False) -> Stats:
        """All runs of the model on one GPU (the reference's Model::runSimulation)."""
        st = Stats()
        self._check(self.lib.psim_model_run(self.handle, device, int(seed), steps_per_launch, int(verbose), C.byref(st)))
        self.runs_done = self.info.num_runs
        return st

    def run_devices(self, devices, seed: int = 1, steps_per_launch: int = 0, verbose: bool = False) -> Stats:
        """All runs of the model with the phonons shared between several GPUs of this process."""
        st = Stats()
        arr = (C.c_int * len(devices))(*devices)
        self._check(self.lib.psim_model_run_devices(self.handle, arr, len(devices), int(seed), steps_per_launch, int(verbose), C.byref(st)))
        self.runs_done = self.info.num_runs
        return st

    def results(self, run_id: Optional[int] = 0, traces: bool = True):
        i = self.info
        S, R = i.num_sensors, i.recorded_steps
        six = np.zeros((S, 6))
        temps = np.zeros((S, R)) if traces else None
        fluxes = np.zeros((S, R, 2)) if traces else None
        rid = 0xFFFFFFFFFFFFFFFF if run_id is None else run_id
        self._check(self.lib.psim_model_results(self.handle, rid, _ptr(six), _ptr(temps), _ptr(fluxes)))
        return six, temps, fluxes

    @property
    def energy_per_phonon(self) -> float:
        return self.lib.psim_model_energy_per_phonon(self.handle)

    def export(self, model_path: str, seconds: float):
        self._check(self.lib.psim_model_export(self.handle, os.fsencode(model_path), float(seconds)))

    def export_text(self, filename: str, seconds: float, when: str = "") -> str:
        n = C.c_size_t(0)
        self._check(self.lib.psim_model_export_text(self.handle, filename.encode(), seconds, when.encode(), None, C.byref(n)))
        buf = C.create_string_buffer(n.value + 1)
        n2 = C.c_size_t(n.value + 1)
        self._check(self.lib.psim_model_export_text(self.handle, filename.encode(), seconds, when.encode(), buf, C.byref(n2)))
        return buf.value.decode()


class GpuSimulator:
    """One psim_gpu handle: the B200 replacement of the reference's ModelSimulator for one shard of the phonons."""

    def __init__(self, desc, device: int = -1):
        self.lib = load_library()
        self.handle = C.c_void_p()
        rc = self.lib.psim_gpu_create(desc, device, C.byref(self.handle))
        if rc:
            raise PsimError(rc, self.lib.psim_gpu_last_error(None).decode())
        d = desc.contents if hasattr(desc, "contents") else desc
        self.num_sensors = d.num_sensors
        self.num_cells = d.num_cells
        self.measurement_steps = d.measurement_steps
        self.recorded_steps = d.measurement_steps - d.step_adjustment

    def close(self):
        if getattr(self, "handle", None):
            self.lib.psim_gpu_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc:
            raise PsimError(rc, self.lib.psim_gpu_last_error(self.handle).decode())

    def set_option(self, name: str, value: int):
        self._check(self.lib.psim_gpu_set_option(self.handle, name.encode(), int(value)))

    def set_sources(self, sources, n: int, seed: int, shard: int = 0, num_shards: int = 1):
        self._check(self.lib.psim_gpu_set_sources(self.handle, sources, n, int(seed), shard, num_shards))

    def run(self):
        self._check(self.lib.psim_gpu_run(self.handle))

    def run_steps(self, begin: int, end: int, stream: int = 0):
        self._check(self.lib.psim_gpu_run_steps(self.handle, begin, end, C.c_void_p(stream) if stream else None))

    def next_window(self, begin: int) -> int:
        """End of the launch window the library would start at measurement step `begin`."""
        end = C.c_uint32()
        self._check(self.lib.psim_gpu_next_window(self.handle, begin, C.byref(end)))
        return end.value

    def synchronize(self):
        self._check(self.lib.psim_gpu_synchronize(self.handle))

    def reset(self):
        self._check(self.lib.psim_gpu_reset(self.handle))

    def tallies(self, fixed: bool = False):
        S, R = self.num_sensors, self.recorded_steps
        e = np.zeros((S, R), dtype=np.int32)
        f = np.zeros((S, R, 2), dtype=np.float64)
        fx = np.zeros((S, R, 2), dtype=np.int64) if fixed else None
        self._check(self.lib.psim_gpu_get_tallies(self.handle, _ptr(e), _ptr(f), _ptr(fx)))
        return (e, f, fx) if fixed else (e, f)

    def tally_buffers(self):
        e, f = C.c_void_p(), C.c_void_p()
        r, s = C.c_uint32(), C.c_uint32()
        self._check(self.lib.psim_gpu_tally_buffers(self.handle, C.byref(e), C.byref(f), C.byref(r), C.byref(s)))
        return e.value, f.value, r.value, s.value

    def alive(self) -> int:
        n = C.c_uint64()
        self._check(self.lib.psim_gpu_alive(self.handle, C.byref(n)))
        return n.value

    def cell_histogram(self) -> np.ndarray:
        h = np.zeros(self.num_cells, dtype=np.uint64)
        self._check(self.lib.psim_gpu_cell_histogram(self.handle, _ptr(h)))
        return h

    def stats(self) -> Stats:
        st = Stats()
        self._check(self.lib.psim_gpu_get_stats(self.handle, C.byref(st)))
        return st

    def probe_sample(self, table: int, u1: np.ndarray, u2: np.ndarray):
        u1 = np.ascontiguousarray(u1, dtype=np.float32)
        u2 = np.ascontiguousarray(u2, dtype=np.float32)
        b = np.zeros(u1.size, dtype=np.uint32)
        t = np.zeros(u1.size, dtype=np.uint32)
        plain = np.zeros(u1.size, dtype=np.uint32)
        self._check(self.lib.psim_gpu_probe_sample(self.handle, table, _ptr(u1), _ptr(u2), u1.size, _ptr(b), _ptr(t), _ptr(plain)))
        return b, t, plain

    def probe_flight(self, cell: np.ndarray, state: np.ndarray) -> np.ndarray:
        cell = np.ascontiguousarray(cell, dtype=np.uint32)
        state = np.ascontiguousarray(state, dtype=np.float32)
        out = np.zeros((cell.size, 6), dtype=np.float32)
        self._check(self.lib.psim_gpu_probe_flight(self.handle, _ptr(cell), _ptr(state), cell.size, _ptr(out)))
        return out

    def probe_rates(self, sensor: int, omega: np.ndarray, ta: np.ndarray) -> np.ndarray:
        omega = np.ascontiguousarray(omega, dtype=np.float64)
        ta = np.ascontiguousarray(ta, dtype=np.uint32)
        out = np.zeros((omega.size, 3))
        self._check(self.lib.psim_gpu_probe_rates(self.handle, sensor, _ptr(omega), _ptr(ta), omega.size, _ptr(out)))
        return out
