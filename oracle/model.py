"""TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT.

numpy restatement of the reference's host-side arithmetic around the particle loop, independent of the C++ host
layer under psim_b200/csrc/host (which it is used to check): material tables, geometry set-up, energy
bookkeeping, phonons per source, tally interpretation.  Each function cites the reference file:line it follows.
The per-phonon loop itself is restated in C (oracle/sim.c) and called from `OracleModel.run`.

Pinned against the unmodified reference by tests/test_oracle.py (tables / energies to 1e-12, tallies to 3 sigma,
fixtures in tests/golden/).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle_sim.so")

HBAR = 1.054517e-34          # material.cpp:12
BOLTZ = 1.38065e-23          # material.cpp:13
BINS = 1000                  # material.h:14
GEOEPS = 2.220446049250313e-16 * 1e9   # utils.h:10
TEMP_INTERVAL = float(np.float32(0.1))  # model.cpp:25 (a float literal widened to double)
SS_STEPS_PERCENT = 0.1       # model.cpp:22


# ----------------------------------------------------------------------------------------------- material
class OracleMaterial:
    """Material::Material (material.cpp:20-51) and the table builders (material.cpp:101-204,241-246)."""

    def __init__(self, mid, m, full):
        d, r = m["d_data"], m["r_data"]
        self.id, self.name, self.full = mid, m["name"], full
        self.la, self.ta = np.array(d["la_data"], float), np.array(d["ta_data"], float)
        self.w_max_la, self.w_max_ta = float(d["max_freq_la"]), float(d["max_freq_ta"])
        self.b_l, self.b_tn, self.b_tu, self.b_i, self.w = (float(r[k]) for k in ("b_l", "b_tn", "b_tu", "b_i", "w"))
        self.freq_width = max(self.w_max_la, self.w_max_ta) / BINS
        self.freq = (2 * np.arange(BINS) + 1) * self.freq_width / 2.0
        with np.errstate(invalid="ignore"):
            k_la, k_ta = self._k(self.freq, self.la), self._k(self.freq, self.ta)
            self.vel_la = 2.0 * self.la[0] * k_la + self.la[1]
            self.dens_la = k_la ** 2 / 2.0 / np.pi ** 2 / self.vel_la
            gv_ta = 2.0 * self.ta[0] * k_ta + self.ta[1]
            ok = ~np.isnan(gv_ta)
            self.vel_ta = np.where(ok, gv_ta, 0.0)
            self.dens_ta = np.where(ok, k_ta ** 2 / np.pi ** 2 / np.where(ok, gv_ta, 1.0), 0.0)
        self.temps = None
        self._cache = {}

    @staticmethod
    def _k(freq, c):  # Material::getK, material.cpp:86-91
        d = c[1] ** 2 - 4.0 * c[0] * (c[2] - freq)
        a = (-c[1] - np.sqrt(d)) / (2.0 * c[0])
        b = (-c[1] + np.sqrt(d)) / (2.0 * c[0])
        return np.where(a < b, a, b)

    def set_grid(self, low, high):  # Material::initializeTables grid, material.cpp:101-109
        steps = int((high - low) / TEMP_INTERVAL)
        self.temps = np.append(low + TEMP_INTERVAL * np.arange(steps, dtype=float), high)
        self._cache = {}

    def temp_index(self, temp):  # getTempIndex, material.cpp:241-246
        return min(int(np.searchsorted(self.temps, temp, side="left")), len(self.temps) - 1)

    def relax_rates(self, temp, freq, ta):  # material.cpp:54-57,207-239 (vectorised over freq)
        freq = np.asarray(freq, float)
        if not ta:
            n = self.b_l * freq * freq * temp ** 3
            u = n.copy()
        else:
            n = np.where(freq < self.w, self.b_tn * freq * temp ** 4, 0.0)
            with np.errstate(over="ignore"):
                u = np.where(freq >= self.w, self.b_tu * freq * freq / np.sinh(HBAR * freq / (temp * BOLTZ)), 0.0)
        return n, u, self.b_i * freq ** 4

    def phonon_dist(self, temp, ta):  # material.cpp:184-204
        dens = self.dens_ta if ta else self.dens_la
        c = HBAR / (BOLTZ * temp)
        with np.errstate(over="ignore", invalid="ignore"):
            d = self.freq * HBAR / np.expm1(c * self.freq) * self.freq_width * dens
            if not self.full:
                d = d * (c * self.freq * np.exp(c * self.freq) / (np.expm1(c * self.freq) * temp))
        return d

    def table(self, kind, temp):
        """kind 0 base / 1 emit / 2 scatter -> (table[1000][2], sum, grid index)  (material.cpp:111-180)"""
        idx = self.temp_index(temp)
        key = (kind, idx)
        if key not in self._cache:
            t = self.temps[idx]
            la, ta = self.phonon_dist(t, False), self.phonon_dist(t, True)
            if kind == 1:
                la, ta = la * self.vel_la, ta * self.vel_ta
            elif kind == 2:
                la = la * sum(self.relax_rates(t, self.freq, False))
                ta = ta * sum(self.relax_rates(t, self.freq, True))
            total = float(np.add.reduce(la) + np.add.reduce(ta))
            tab = np.empty((BINS, 2))
            tab[:, 0] = np.cumsum((la + ta) / total)
            with np.errstate(invalid="ignore", divide="ignore"):
                tab[:, 1] = la / (la + ta)
            self._cache[key] = (tab, total, idx)
        return self._cache[key]


# ------------------------------------------------------------------------------------------------ geometry
def _on_line(p1, p2, q):  # isPointOnLine, geometry.cpp:295-298 (p1, p2: [...,2]; q: [...,2])
    return np.abs((p2[..., 0] - p1[..., 0]) * (q[..., 1] - p1[..., 1]) - (q[..., 0] - p1[..., 0]) * (p2[..., 1] - p1[..., 1])) < GEOEPS


def _contains(a1, a2, b1, b2):  # Line::contains(Line), geometry.cpp:75-86 : segment a contains segment b
    la = np.hypot(a2[..., 0] - a1[..., 0], a2[..., 1] - a1[..., 1])
    lb = np.hypot(b2[..., 0] - b1[..., 0], b2[..., 1] - b1[..., 1])
    ok = _on_line(a1, a2, b1) & _on_line(a1, a2, b2) & (la >= lb)
    for ax in (0, 1):
        amax, amin = np.maximum(a1[..., ax], a2[..., ax]), np.minimum(a1[..., ax], a2[..., ax])
        bmax, bmin = np.maximum(b1[..., ax], b2[..., ax]), np.minimum(b1[..., ax], b2[..., ax])
        ok &= (amax >= bmax - GEOEPS) & (amin <= bmin + GEOEPS)
    return ok


class OracleModel:
    def __init__(self, model: dict):
        st = model["settings"]
        self.num_runs = int(st.get("num_runs", 1))
        self.M = int(st["num_measurements"])
        self.num_phonons = int(st["num_phonons"])
        self.sim_time = float(st["sim_time"])
        self.t_eq = float(st["t_eq"])
        self.phasor = bool(st["phasor_sim"])
        self.sim_type = int(st["sim_type"]) if int(st["sim_type"]) in (1, 2) else 0
        self.step_interval = int(st.get("step_interval", 0))
        # Model::setSimulationType / addSensor, model.cpp:46-70,89-91
        self.start_step = int(self.M - self.M * SS_STEPS_PERCENT) if self.sim_type != 0 else 0
        self.step_adjustment = int(self.M - self.M * SS_STEPS_PERCENT) if self.sim_type == 0 else 0
        self.R = int(self.M * SS_STEPS_PERCENT) if self.sim_type == 0 else self.M
        full = self.t_eq == 0.0
        self.materials = [OracleMaterial(i, m, full) for i, m in enumerate(model["materials"])]
        mat_id = {m.name: m.id for m in self.materials}
        self.sensor_ids = np.array([s["id"] for s in model["sensors"]], dtype=np.int64)
        sidx = {int(s["id"]): i for i, s in enumerate(model["sensors"])}
        self.sensor_material = np.array([mat_id[s["material"]] for s in model["sensors"]], dtype=np.int32)
        self.t_init = np.array([float(s["t_init"]) for s in model["sensors"]])
        self.t_steady = self.t_init.copy()
        tri = np.array([[[c["triangle"][p]["x"], c["triangle"][p]["y"]] for p in ("p1", "p2", "p3")] for c in model["cells"]], float)
        self.tri = tri
        self.cell_sensor = np.array([sidx[int(c["sensorID"])] for c in model["cells"]], dtype=np.int32)
        self.cell_spec = np.array([float(c["specularity"]) for c in model["cells"]])
        # Triangle::isClockwise, geometry.cpp:217-222
        nxt = np.roll(tri, -1, axis=1)
        self.norm_sign = np.where(((nxt[..., 0] - tri[..., 0]) * (nxt[..., 1] + tri[..., 1])).sum(axis=1) >= 0.0, 1, -1).astype(np.int32)
        # Triangle::area (Heron), geometry.cpp:250-258
        ln = np.hypot(nxt[..., 0] - tri[..., 0], nxt[..., 1] - tri[..., 1])
        p = ln.sum(axis=1) / 2.0
        self.cell_area = np.sqrt(p * (p - ln[:, 0]) * (p - ln[:, 1]) * (p - ln[:, 2]))
        self.sensor_area = np.zeros(len(self.sensor_ids))
        np.add.at(self.sensor_area, self.cell_sensor, self.cell_area)
        self._find_transitions()
        self._attach_emitters(model["emit_surfaces"])

    # Model::addCell -> Cell::findTransitionSurface (model.cpp:98-113, cell.cpp:81-98,126-132)
    def _find_transitions(self):
        tri = self.tri
        n = len(tri)
        e1, e2 = tri, np.roll(tri, -1, axis=1)  # edge k: vertex k -> k+1
        xmin, xmax = tri[..., 0].min(axis=1), tri[..., 0].max(axis=1)
        ymin, ymax = tri[..., 1].min(axis=1), tri[..., 1].max(axis=1)
        subs = [[[] for _ in range(3)] for _ in range(n)]  # (target, x1, y1, x2, y2)
        for i in range(1, n):
            near = np.nonzero((xmin[:i] <= xmax[i] + 1e-6) & (xmax[:i] >= xmin[i] - 1e-6) & (ymin[:i] <= ymax[i] + 1e-6) & (ymax[:i] >= ymin[i] - 1e-6))[0]
            if near.size == 0:
                continue
            for j in near:  # existing cells in model order
                for a in range(3):       # l1: incoming cell's lines
                    for b in range(3):   # l2: existing cell's lines
                        if _contains(e1[i, a], e2[i, a], e1[j, b], e2[j, b]):
                            line = (e1[j, b], e2[j, b])
                        elif _contains(e1[j, b], e2[j, b], e1[i, a], e2[i, a]):
                            line = (e1[i, a], e2[i, a])
                        else:
                            continue
                        rec = (line[0][0], line[0][1], line[1][0], line[1][1])
                        subs[i][a].append((int(j),) + rec)
                        subs[j][b].append((int(i),) + rec)
        self.transitions = subs

    # Model::setEmitSurface -> Cell::setEmitSurface -> CompositeSurface::addEmitSurface (model.cpp:125-138,
    # cell.cpp:29-35, compositeSurface.cpp:20-35)
    def _attach_emitters(self, surfaces):
        tri = self.tri
        e1, e2 = tri, np.roll(tri, -1, axis=1)
        self.emitters = [[[] for _ in range(3)] for _ in range(len(tri))]
        self.emit_list = []
        for s in surfaces:
            p1 = np.array([s["p1"]["x"], s["p1"]["y"]], float)
            p2 = np.array([s["p2"]["x"], s["p2"]["y"]], float)
            ok = _contains(e1, e2, p1[None, None, :], p2[None, None, :])  # [cells, 3]
            hits = np.argwhere(ok)
            if hits.size == 0:
                raise RuntimeError("Unable to add emitting surface.")
            c, k = (int(v) for v in hits[0])  # first cell in model order, first edge of it
            rec = dict(cell=c, edge=k, p1=p1, p2=p2, temp=float(s["temp"]), duration=float(s["duration"]), start=float(s["start_time"]),
                       length=float(np.hypot(*(p2 - p1))))
            self.emitters[c][k].append(rec)
            self.emit_list.append(rec)

    def init_temp(self):  # SteadyState: t_steady_; others: t_init_ (sensorController.h:65-67,83-85,100-102)
        return self.t_steady if self.sim_type == 0 else self.t_init

    # Model::setTemperatureBounds + initializeMaterialTables + SensorController::updateTables
    # (model.cpp:203-227, sensorController.cpp:38-50)
    def prepare(self):
        temps = list(self.init_temp()[self.cell_sensor]) + [e["temp"] for e in self.emit_list]
        lo, hi = min(temps), max(temps)
        bound = 1000.0 if self.phasor else 10.0
        self.lb, self.ub = max(lo - bound, 0.0), hi + bound
        for m in self.materials:
            m.set_grid(lo, hi)
        self.heat_capacity = np.array([self.materials[self.sensor_material[s]].table(0, self.t_init[s])[1] for s in range(len(self.t_init))])
        self.refresh()

    # Model::getTotalInitialEnergy (model.cpp:196-201) with Cell::getInitEnergy / getEmitEnergy (cell.cpp:38-63)
    def cell_energies(self):
        it = self.init_temp()[self.cell_sensor]
        init = self.cell_area * self.heat_capacity[self.cell_sensor]
        if self.t_eq != 0.0:
            init = init * np.abs(it - self.t_eq)
        emit = np.zeros(len(self.tri))
        for c in range(len(self.tri)):
            mat = self.materials[self.sensor_material[self.cell_sensor[c]]]
            for k in range(3):
                for e in self.emitters[c][k]:
                    en = e["length"] * e["duration"] * mat.table(1, e["temp"])[1] / 4.0
                    emit[c] += en if self.t_eq == 0.0 else en * abs(e["temp"] - self.t_eq)
        return init, emit

    def total_energy(self):
        init, emit = self.cell_energies()
        total = 0.0
        for a, b in zip(init, emit):
            total += a + b
        return total

    def refresh(self):  # the `refresh` lambda, model.cpp:148-153
        self.eff_energy = self.total_energy() / float(self.num_phonons)

    # ModelSimulator::initPhononBuilders, modelSimulator.cpp:43-85 (seeded instead of random_device)
    def sources(self, seed):
        rng = np.random.default_rng(seed)
        init, _ = self.cell_energies()
        out = []
        sub_index = self._sub_index()

        def phonons(energy):
            frac, whole = np.modf(energy / self.eff_energy)
            return int(whole) + (1 if rng.random() < frac else 0)

        it = self.init_temp()
        for c in range(len(self.tri)):
            s = self.cell_sensor[c]
            n = phonons(init[c])
            if n > 0:
                out.append((0, c, -1, 1 if it[s] > self.t_eq else -1, n))
            mat = self.materials[self.sensor_material[s]]
            for k in range(3):
                for e in self.emitters[c][k]:
                    factor = mat.table(1, e["temp"])[1] * e["duration"] * e["length"] / 4.0
                    en = factor if self.t_eq == 0.0 else factor * abs(self.t_eq - e["temp"])
                    out.append((1, c, sub_index[id(e)], 1 if e["temp"] > self.t_eq else -1, phonons(en)))
        return np.array(out, dtype=np.int64).reshape(-1, 5)

    def _flat_subs(self):
        """Sub-surfaces sorted by (cell, edge), transitions before emitters, insertion order kept."""
        kind, cell, edge, target, line, normal, table, window = [], [], [], [], [], [], [], []
        tables, table_ids = [], {}

        def tid(mat, k, temp):
            tab, _, idx = mat.table(k, temp)
            key = (mat.id, k, idx)
            if key not in table_ids:
                table_ids[key] = len(tables)
                tables.append(tab)
            return table_ids[key]

        sensor_base = [tid(self.materials[self.sensor_material[s]], 0, self.t_init[s]) for s in range(len(self.t_init))]
        sensor_scatter = [tid(self.materials[self.sensor_material[s]], 2, self.t_init[s]) for s in range(len(self.t_init))]
        index = {}
        tri = self.tri
        for c in range(len(tri)):
            ns = self.norm_sign[c]
            mat = self.materials[self.sensor_material[self.cell_sensor[c]]]
            for k in range(3):
                for (tgt, x1, y1, x2, y2) in self.transitions[c][k]:
                    ln = np.hypot(x2 - x1, y2 - y1)
                    kind.append(1); cell.append(c); edge.append(k); target.append(tgt); line.append((x1, y1, x2, y2))
                    normal.append((ns * (y2 - y1) / ln, -ns * (x2 - x1) / ln))  # Surface ctor + Line::normal (surface.cpp:11-16, geometry.cpp:97-100)
                    table.append(0); window.append((0.0, 0.0, 0.0))
                a, b = tri[c, k], tri[c, (k + 1) % 3]
                ln = np.hypot(*(b - a))
                main_n = (ns * (b[1] - a[1]) / ln, -ns * (b[0] - a[0]) / ln)
                for e in self.emitters[c][k]:
                    index[id(e)] = len(kind)
                    kind.append(2); cell.append(c); edge.append(k); target.append(-1)
                    line.append((e["p1"][0], e["p1"][1], e["p2"][0], e["p2"][1]))
                    normal.append(main_n)  # setNormal(main_surface_.getNormal()), compositeSurface.cpp:31
                    table.append(tid(mat, 1, e["temp"])); window.append((e["temp"], e["start"], e["duration"]))
        arr = lambda x, t: np.ascontiguousarray(np.array(x, dtype=t).reshape(len(x), -1) if len(x) else np.zeros((0, 1), dtype=t))
        return dict(kind=arr(kind, np.int32), cell=arr(cell, np.int32), edge=arr(edge, np.int32), target=arr(target, np.int32),
                    line=arr(line, np.float64), normal=arr(normal, np.float64), table=arr(table, np.int32), window=arr(window, np.float64),
                    tables=np.ascontiguousarray(np.array(tables)), sensor_base=np.array(sensor_base, dtype=np.int32),
                    sensor_scatter=np.array(sensor_scatter, dtype=np.int32), index=index)

    def _sub_index(self):
        return self._flat_subs()["index"]

    # ------------------------------------------------------------------------------------------- hot path (C)
    def run(self, seed: int, threads: int = 0):
        """One run of the restated particle loop -> (energy[S][R] int64, flux[S][R][2], drift_steps, loop_iters)."""
        lib = load_sim()
        f = self._flat_subs()
        src = np.ascontiguousarray(self.sources(seed))
        S, R = len(self.t_init), self.R
        energy = np.zeros((S, R), dtype=np.int64)
        flux = np.zeros((S, R, 2))
        counters = np.zeros(2, dtype=np.int64)
        cell_xy = np.ascontiguousarray(self.tri.reshape(len(self.tri), 6))
        mat_consts = np.ascontiguousarray(np.array([[m.b_l, m.b_tn, m.b_tu, m.b_i, m.w, m.w_max_la, m.w_max_ta, m.freq_width] for m in self.materials]))
        mat_arrays = np.ascontiguousarray(np.array([[m.freq, m.vel_la, m.vel_ta] for m in self.materials]))
        # getSteadyTemp(): t_steady_ (== t_init during a run since MAX_ITERS = 1, model.cpp:11); transient: t_init
        sensor_temp = np.ascontiguousarray(self.t_init if self.sim_type == 2 else self.t_steady)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = lib.oracle_run(len(self.tri), p(cell_xy), p(self.cell_sensor), p(np.ascontiguousarray(self.cell_spec)), p(self.norm_sign),
                            len(f["kind"]) if f["kind"].size and f["kind"].shape[1] and len(f["index"]) + sum(len(t) for c in self.transitions for t in c) else 0,
                            p(f["kind"]), p(f["cell"]), p(f["edge"]), p(f["target"]), p(f["line"]), p(f["normal"]), p(f["table"]), p(f["window"]),
                            S, p(self.sensor_material), p(sensor_temp), p(f["sensor_base"]), p(f["sensor_scatter"]),
                            len(self.materials), p(mat_consts), p(mat_arrays), len(f["tables"]), p(f["tables"]),
                            self.M, self.step_adjustment, C.c_double(self.sim_time), int(self.t_eq == 0.0), int(self.phasor),
                            len(src), p(src), C.c_uint64(seed), threads, p(energy), p(flux), p(counters))
        if rc:
            raise RuntimeError("oracle_run failed")
        self.inc_energy, self.inc_flux = energy, flux
        return energy, flux, int(counters[0]), int(counters[1])

    # ------------------------------------------------------------------------------------- interpretation
    # SensorInterpreter::findTemperature, sensorInterpreter.cpp:80-112
    def find_temperature(self, s, start=0):
        e = self.eff_energy * self.inc_energy[s, start:].astype(float)
        if self.t_eq != 0.0:
            return e / (self.sensor_area[s] * self.heat_capacity[s]) + self.t_eq
        mat = self.materials[self.sensor_material[s]]
        out = np.zeros(len(e))
        for i, en in enumerate(e):
            temp, ub, lb, it = 0.0, self.ub, self.lb, 0
            while ub - lb >= 1e-4:
                it += 1
                if it == 40:
                    break
                temp = (ub + lb) / 2.0
                if mat.table(0, temp)[1] * self.sensor_area[s] - en < 0.0:
                    lb = temp
                else:
                    ub = temp
            out[i] = temp
        return out

    # Model::resetRequired side effects + refresh + storeResults (model.cpp:163-177,250-272; sensorInterpreter.cpp:19-66)
    def finish_run(self):
        S = len(self.t_init)
        if self.sim_type != 2:
            for s in range(S):
                temps = self.find_temperature(s, self.start_step)
                self.t_steady[s] = temps.sum() / float(self.R - self.start_step) if self.sensor_area[s] != 0.0 else 0.0
        self.refresh()
        six = np.zeros((S, 6))
        temps_all = np.zeros((S, self.R))
        flux_all = np.zeros((S, self.R, 2))
        it = self.init_temp()
        for s in range(S):
            t = self.find_temperature(s, 0)
            t[0] = it[s]
            fl = self.inc_flux[s] * (self.eff_energy / self.sensor_area[s])
            temps_all[s], flux_all[s] = t, fl
            for col, data in ((0, t), (2, fl[:, 0]), (4, fl[:, 1])):
                avg = data.sum() / len(data)
                six[s, col] = avg
                six[s, col + 1] = np.sqrt(((avg - data) ** 2).sum() / len(data)) / np.sqrt(len(data))
        order = np.argsort(self.sensor_ids, kind="stable")  # sortMeasurements, outputManager.cpp:120-122
        return six[order], temps_all[order], flux_all[order]


_sim = None


def load_sim():
    global _sim
    if _sim is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "sim.c")):
            subprocess.run(["make", "-C", HERE, "oracle"], check=True, stdout=subprocess.DEVNULL)
        lib = C.CDLL(LIB)
        lib.oracle_run.restype = C.c_int
        vp = C.c_void_p
        lib.oracle_run.argtypes = ([C.c_int, vp, vp, vp, vp] + [C.c_int] + [vp] * 8 + [C.c_int, vp, vp, vp, vp] +
                                   [C.c_int, vp, vp, C.c_int, vp] + [C.c_int64, C.c_int64, C.c_double, C.c_int, C.c_int] +
                                   [C.c_int, vp, C.c_uint64, C.c_int] + [vp, vp, vp])
        lib.oracle_flight.restype = C.c_int
        lib.oracle_flight.argtypes = [vp, vp, C.c_double, vp]
        _sim = lib
    return _sim


def oracle_flight(tri, state, horizon=1.0e4):
    """Reference geometry of one free flight + specular reflection (oracle/sim.c:oracle_flight)."""
    lib = load_sim()
    tri = np.ascontiguousarray(tri, dtype=np.float64)
    state = np.ascontiguousarray(state, dtype=np.float64)
    out = np.zeros(6)
    lib.oracle_flight(tri.ctypes.data_as(C.c_void_p), state.ctypes.data_as(C.c_void_p), float(horizon), out.ctypes.data_as(C.c_void_p))
    return out
