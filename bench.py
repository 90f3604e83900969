#!/usr/bin/env python
"""Benchmark of the phonon Monte Carlo particle loop (BASELINE.json: phonon drift-steps/sec at 1/2/4/8 B200 + roofline;
wall-clock per model).

Workload (config.workload): BASELINE.json configs[4], the synthetic 100-cell Si/Ge structure with 1e8 deviational phonons
IN TOTAL, sharded over the N GPUs of the job (strong scaling: rank r owns the phonon ids == r mod N), 1000 measurement
steps over 1 ns.  One bench "step" = one complete simulation of that model: every measurement interval of every phonon
(emission, drift, scattering, surfaces, cell transitions, tallies) plus the NCCL all-reduce of the sensor tallies.
1 drift-step = one live phonon advanced across one measurement interval.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference ...                         the reference's own CPU code on the host cores

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over
ranks.  Besides the headline the line carries
  strong      the 64-bit checksums that must be identical at N = 1, 2, 4, 8: reduced int32 / int64 tallies, emitted counts,
              phonons per cell at mid-run - for the bench job and for the 6174-cell kinked wire (31 MB all-reduce)
  weak        the same job with 1e8 phonons PER GPU (N > 1 only; at N = 1 it is the headline)
  models      wall-clock per shipped model at its full phonon count (N = 1), next to the reference's own headers
  roofline    what bounds the kernel: instruction issue (warp instructions per drift-step from the committed ncu captures,
              used only while profiles/r02_ncu_summary.json carries the hash of the device sources it was taken on and
              the job that runs has the captured job's launches and segments per drift-step) and the measured DRAM traffic
              against the HBM peak
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from psim_b200 import configs  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "psim_ref")
METRIC = "phonon drift-steps/sec"
NCU_SUMMARY = os.path.join(ROOT, "profiles", "r02_ncu_summary.json")
WORKLOAD = "synthetic 100-cell Si/Ge 2D structure (BASELINE.json configs[4]), steady-state deviational"
CHECKSUM_SEED = 4242
# the reference's own result-file headers (psim_python/json/results/ss_*.txt:1, "Time Taken ...[s]")
REFERENCE_HEADER_SECONDS = {"linear_demo": 18.1, "linear_sides_demo_ss": 32.0, "kinked_demo_120_35_spec": 212.7}


def workload_model(num_phonons: int) -> dict:
    return configs.si_ge_grid(num_phonons=num_phonons).to_dict()


# what decides the work of the kernels: the device code, its launcher and the builder of the device images (the host layer
# under csrc/host/ - loader, run epilogue, exporter - only feeds them a model description)
DEVICE_SOURCES = ("device_core.cuh", "device_types.h", "flatten.cpp", "flatten.h", "kernels.cuh", "psim_gpu.cu")


def _sha16(paths, base) -> str:
    h = hashlib.sha256()
    for p in paths:
        h.update(os.path.relpath(p, base).encode())
        h.update(open(p, "rb").read())
    return h.hexdigest()[:16]


def csrc_sha16() -> str:
    """Hash of every source of the library (kernels, C ABI, host layer)."""
    base = os.path.join(ROOT, "psim_b200", "csrc")
    return _sha16([os.path.join(d, f) for d, _, files in sorted(os.walk(base)) for f in sorted(files)], base)


def device_sha16() -> str:
    """Hash of DEVICE_SOURCES: profiles/*_ncu_summary.json records the value its counters were taken on, and the ncu-derived
    constants are used only while it matches the tree that is running (and the job that runs matches the captured one,
    `ncu_constants`)."""
    base = os.path.join(ROOT, "psim_b200", "csrc")
    return _sha16([os.path.join(base, f) for f in DEVICE_SOURCES], base)


def ncu_constants(launches=None, segments_per_drift_step=None, phonons_total=None):
    """profiles/r02_ncu_summary.json if it was captured on THESE device sources and describes THIS job - the same phonons in
    total, the same number of launches, the same flight segments per drift-step to 0.5 % (counted live by the run: a change
    anywhere that altered the kernels' work would show there) - else None (never a stale constant)."""
    try:
        d = json.load(open(NCU_SUMMARY))
        if d.get("device_sha16") != device_sha16():
            return None
        if phonons_total is not None and int(phonons_total) != int(d["phonons_per_gpu"]):  # captured on one GPU
            return None
        j = d["bench_job"]
        if launches is not None and int(launches) != int(j["launches"]):
            return None
        captured = j["segments_per_job"] / j["drift_steps_per_job"]
        if segments_per_drift_step is not None and abs(segments_per_drift_step / captured - 1.0) > 5e-3:
            return None
        return d
    except Exception:
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(path))
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


def checksum64(*arrays) -> str:
    h = hashlib.blake2b(digest_size=8)
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


class ClockSampler:
    """nvidia-smi SM clock + throttle reasons while the timed region runs (rank 0 only)."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None

    def _loop(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=10)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------- CPU reference
def oracle_drift_steps_per_phonon(model: dict, phonons: int = 100_000) -> float | None:
    """Drift-steps per phonon of `model`, COUNTED on the CPU by the restatement of the reference's event loop (oracle/sim.c
    counts one per measurement event and one per exit) on a sample of `phonons` phonons.  The unmodified
    reference binary has no such counter; its rate is its phonons per second times this figure."""
    try:
        from oracle.model import OracleModel
        om = OracleModel(configs.with_settings(model, num_phonons=phonons))
        om.prepare()
        _, _, steps, _ = om.run(1, threads=min(os.cpu_count() or 1, 16))
        return float(steps) / float(phonons)
    except Exception:
        return None


def reference_cpu_run(model: dict, phonons_per_proc: int, procs: int, drift_steps_per_phonon: float | None):
    """Runs the reference's own CPU implementation (oracle/_ref/psim_ref: unmodified reference sources + our
    driver main) as `procs` independent processes (no TBB in this image, so std::execution::par is serial; phonons
    are independent, so P processes of n/P phonons are the reference's parallel path).  Wall = slowest process."""
    if not os.path.exists(REF_BIN):
        return oracle_cpu_run(model, phonons_per_proc * procs, procs)
    with tempfile.TemporaryDirectory() as tmp:
        path = configs.save(configs.with_settings(model, num_phonons=phonons_per_proc), os.path.join(tmp, "m.json"))
        t0 = time.perf_counter()
        ps = [subprocess.Popen([REF_BIN, "run", path, os.path.join(tmp, f"o{i}")], stdout=subprocess.DEVNULL) for i in range(procs)]
        rcs = [p.wait() for p in ps]
        wall = time.perf_counter() - t0
        if any(rcs):
            return None
        inner = max(json.load(open(os.path.join(tmp, f"o{i}.meta.json")))["seconds"] for i in range(procs))
    total = phonons_per_proc * procs
    out = {"phonons": total, "seconds": inner, "wall_with_load": wall, "kind": "reference"}
    if drift_steps_per_phonon:
        out["drift_steps_per_s"] = total * drift_steps_per_phonon / inner
    return out


def oracle_cpu_run(model: dict, phonons: int, threads: int):
    """Fallback when the reference binary is not on the box: the plain-C restatement under oracle/ (OpenMP), which counts
    its own drift-steps."""
    try:
        from oracle.model import OracleModel
        om = OracleModel(configs.with_settings(model, num_phonons=phonons))
        om.prepare()
        t0 = time.perf_counter()
        _, _, steps, _ = om.run(1, threads=threads)
        secs = time.perf_counter() - t0
    except Exception:
        return None
    return {"phonons": phonons, "seconds": secs, "wall_with_load": secs, "kind": "port", "drift_steps_per_s": steps / secs}


# --------------------------------------------------------------------------------------------------------- ours
_RESULT_FD = None


def claim_stdout():
    """Everything libraries write to stdout (NCCL's version banner, for one) goes to stderr from here on; the result line
    alone is written to the real stdout by emit()."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, line)


def agree_on_cuts(cuts, num_steps: int, world: int, rank: int, device):
    """Every rank must cut a job into the same groups of measurement steps (each group ends with a collective).  The
    library's launch windows depend, through the exactness bound of the tally staging, on the capacity of the rank's own
    pool - which may differ by one phonon between ranks - so rank 0's cuts are broadcast and used by all; a rank whose own
    windows are shorter simply needs two launches for such a group."""
    if world == 1:
        return list(cuts)
    import torch
    import torch.distributed as dist
    t = torch.full((num_steps + 2,), -1, dtype=torch.int64, device=device)
    if rank == 0:
        t[:len(cuts)] = torch.tensor(cuts, dtype=torch.int64)
    dist.broadcast(t, 0)
    return [int(x) for x in t.tolist() if x >= 0]


class Env:
    """Rank / device / process group of this bench process."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the CUDA path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.device = f"cuda:{self.local}"
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
                os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.stream = torch.cuda.Stream(device=self.local)  # the drift kernels are launched on THIS stream, and so are the events
        torch.cuda.set_stream(self.stream)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def sum_over_ranks(self, values):
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]


class _Dev:  # expose the library's tally buffers to torch (for the NCCL all-reduce) without a copy
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


class ShardedJob:
    """One model, its phonons sharded over the ranks by id (rank r owns ids == r mod N): the library handle of this rank,
    torch views of its device-resident tallies, and the groups of measurement steps between tally all-reduces."""

    def __init__(self, env: Env, model_dict: dict, args, options=None):
        from psim_b200 import lib as psim
        self.env, self.args = env, args
        self.model = psim.Model(text=json.dumps(model_dict))
        self.model.prepare()
        info = self.model.info
        self.M, self.S, self.R = info.measurement_steps, info.num_sensors, info.recorded_steps
        self.cells = info.num_cells
        self.first = self.M - self.R  # first step whose measurement is recorded
        self.g = psim.GpuSimulator(self.model.describe(), env.local)
        self.g.set_option("steps_per_launch", args.steps_per_launch)
        for k, v in (options or {}).items():
            if v is not None and v >= 0:
                self.g.set_option(k, v)
        e_ptr, f_ptr, _, _ = self.g.tally_buffers()
        self.t_energy = env.torch.as_tensor(_Dev(e_ptr, (self.R, self.S), "<i4"), device=env.device)
        self.t_flux = env.torch.as_tensor(_Dev(f_ptr, (self.R, self.S, 2), "<i8"), device=env.device)
        self.sources = None

    def close(self):
        self.g.close()
        self.model.close()

    def set_sources(self, seed: int):
        src, n = self.model.sources(seed)
        self.sources = [(src[i].kind, src[i].index, src[i].sign, src[i].count) for i in range(n)]
        self.g.set_sources(src, n, seed, self.env.rank, self.env.world)  # untimed: pool reset, tallies zeroed, birth plan resident in HBM

    def plan_cuts(self, extra_cut: int | None = None):
        """Steps whose measurement is not recorded (steady state: the first 90 %) need no exchange and go to the library
        in one call (it chooses its own launch windows); the recorded steps go in groups that end where the library's
        launch windows end (no extra launch for the exchange; --reduce-every N asks for groups of N steps instead); all
        ranks use rank 0's cuts (agree_on_cuts)."""
        args, g, M, first = self.args, self.g, self.M, self.first
        chunk = max(args.steps_per_launch, args.reduce_every) if args.reduce_every > 0 else 0
        cuts, s = [0], 0
        if first > 1:
            s = first - 1
            cuts.append(s)
        while s < M - 1:
            s = min(s + chunk, M - 1) if chunk > 0 else g.next_window(s)
            cuts.append(s)
        if extra_cut is not None and extra_cut not in cuts:
            cuts = sorted(cuts + [extra_cut])
        return agree_on_cuts(cuts, M, self.env.world, self.env.rank, self.env.device)

    def run(self, cuts, at_cut=None):
        """All measurement steps; the tally all-reduce of a group of steps is issued as soon as the group's launches are
        enqueued."""
        env = self.env
        for s, e in zip(cuts[:-1], cuts[1:]):
            self.g.run_steps(s, e, env.stream.cuda_stream)
            if env.world > 1:
                r0, r1 = max(s + 1 - self.first, 0), e + 1 - self.first  # tally rows completed by steps [s, e)
                if r1 > r0:
                    env.dist.all_reduce(self.t_energy[r0:r1])
                    env.dist.all_reduce(self.t_flux[r0:r1])
            if at_cut is not None:
                at_cut(e)

    def timed_jobs(self, warmup: int, steps: int, seed0: int, sampler=None):
        env = self.env
        times, kernel_ms, drift, events, launches = [], [], [], [], 0
        cuts = None
        for it in range(warmup + steps):
            self.set_sources(seed0 + it)
            cuts = self.plan_cuts()
            timed = it >= warmup
            if timed and it == warmup and sampler:
                sampler.__enter__()
            env.barrier()
            ev0, ev1 = env.torch.cuda.Event(enable_timing=True), env.torch.cuda.Event(enable_timing=True)
            ev0.record(env.stream)
            self.run(cuts)
            ev1.record(env.stream)
            env.barrier()
            st = self.g.stats()
            if timed:
                times.append(ev0.elapsed_time(ev1))
                kernel_ms.append(st.kernel_ms)
                drift.append(st.drift_steps)
                events.append(st.events)
                launches += st.launches
        ms = env.max_over_ranks([np.mean(times)])[0]
        tot = env.sum_over_ranks([np.mean(drift), np.mean(events)])
        return {"ms_per_step": ms, "drift_steps": tot[0], "events": tot[1], "kernel_ms": float(np.mean(kernel_ms)),
                "drift_steps_rank": float(np.mean(drift)), "launches": launches, "cuts": cuts, "stats": self.g.stats().as_dict()}

    def checksum_job(self, seed: int = CHECKSUM_SEED):
        """One job with a fixed seed whose integers must not depend on the number of GPUs: the reduced tallies, the emitted
        counts and the number of phonons per cell half-way through the run."""
        env = self.env
        self.set_sources(seed)
        mid = self.M // 2
        cuts = self.plan_cuts(extra_cut=mid)
        hist = {}

        def at_cut(e):
            if e == mid:
                h = self.g.cell_histogram().astype(np.int64)
                t = env.torch.from_numpy(h).to(env.device)
                if env.world > 1:
                    env.dist.all_reduce(t)
                hist["cells"] = t.cpu().numpy()

        env.barrier()
        self.run(cuts, at_cut)
        env.barrier()
        e = self.t_energy.cpu().numpy()
        f = self.t_flux.cpu().numpy()
        st = self.g.stats()
        tot = env.sum_over_ranks([st.drift_steps, st.shard_phonons])
        counts = np.array([c for (_, _, _, c) in self.sources], dtype=np.uint64)
        return {"seed": seed, "tallies": checksum64(e, f), "energy": checksum64(e), "flux": checksum64(f),
                "emitted_counts": checksum64(counts), "cell_histogram_mid_run": checksum64(hist["cells"]),
                "phonons_emitted": int(counts.sum()), "phonons_alive_mid_run": int(hist["cells"].sum()),
                "drift_steps": int(tot[0]), "allreduce_bytes_per_job": int(e.nbytes + f.nbytes), "cuts": cuts}


def end_to_end(env: Env, model_dict: dict, args, runs: int):
    """End to end through the host API with HOST buffers, as the runs of a multi-run model go (psim_model_run): per run the
    sources / birth plan go host -> device, the GPU runs, the tallies come back device -> host and the reference's run
    epilogue (temperatures / fluxes) is done on the host.  The handle (model image, pool) is created once, as
    psim_model_run does; the first run through it is the warm-up."""
    from psim_b200 import lib as psim
    torch, dist = env.torch, env.dist
    model = psim.Model(text=json.dumps(model_dict))
    model.prepare()
    g2 = psim.GpuSimulator(model.describe(), env.local)  # uploads cells / sensors / tables (host -> device), once
    g2.set_option("steps_per_launch", args.steps_per_launch)
    e2e_ms, detail, h2d, d2h = [], [], 0, 0
    for it in range(1 + runs):
        seed = 2000 + it
        env.barrier()
        t0 = time.perf_counter()
        src, n = model.sources(seed)
        g2.set_sources(src, n, seed, env.rank, env.world)  # birth plan host -> device
        g2.run()
        e, f = g2.tallies()  # device -> host
        if env.world > 1:
            te, tf = torch.from_numpy(e.astype(np.int64)).cuda(), torch.from_numpy(f).cuda()
            dist.all_reduce(te)
            dist.all_reduce(tf)
            e, f = te.cpu().numpy().astype(np.int32), tf.cpu().numpy()
        model.set_tallies(e, f)
        model.finish_run(0)
        model.results(0, traces=False)
        model.next_run()
        model.prepare()
        env.barrier()
        ms = (time.perf_counter() - t0) * 1e3
        st2 = g2.stats()
        detail.append({"ms": round(ms, 2), "kernel_ms": round(st2.kernel_ms, 2), "phonons": st2.total_phonons,
                       "drift_steps": st2.drift_steps, "timed": it > 0})
        if it > 0:
            e2e_ms.append(ms)
        h2d, d2h = st2.plan_bytes, st2.tally_bytes  # sources + birth plan / tallies, counted by the library
    g2.close()
    model.close()
    return env.max_over_ranks([np.mean(e2e_ms)])[0], detail, int(h2d), int(d2h)


def shipped_models():
    from tests import cases
    m = {
        "linear_demo": configs.linear().to_dict(),
        "linear_sides_demo_ss": configs.linear_sides().to_dict(),
        "linear_sides_demo_per": configs.linear_sides(sim_type=1, step_interval=4).to_dict(),
        "linear_sides_demo_trans": configs.linear_sides(sim_type=2, step_interval=4, start_time=0.1, duration=0.15).to_dict(),
    }
    kinked = cases.kinked_model()
    if kinked is not None:
        m["kinked_demo_120_35_spec"] = kinked
        m["kinked_demo_120_35_spec0.5"] = configs.with_specularity(kinked, 0.5)
    m["si_ge_grid_1e8"] = configs.si_ge_grid().to_dict()
    return m


def model_walltimes(device: int):
    """BASELINE.json "wall-clock per model": every shipped configuration at its FULL phonon count, end to end through the
    host API (ms_e2e = the whole psim_model_run: device image and pool set-up, host -> device, kernels, device -> host, run
    epilogue - what the reference's "Time Taken" header measures), the kernel-only time, and the throughput in drift-steps and
    in flight segments.  The process already has its CUDA context (the bench job ran first)."""
    from psim_b200 import lib as psim
    out = []
    with tempfile.TemporaryDirectory() as tmp:
        for name, model in shipped_models().items():
            path = configs.save(model, os.path.join(tmp, "m.json"))
            t0 = time.perf_counter()
            m = psim.Model(path)
            t1 = time.perf_counter()
            load_ms = (t1 - t0) * 1e3          # JSON -> model: parse, mesh validation, neighbour discovery
            st = m.run(device=device, seed=1)  # psim_model_run: device image + pool set-up, H2D, kernels, D2H, run epilogue
            t2 = time.perf_counter()
            first_ms = (t2 - t1) * 1e3
            # Twice: on these boxes a single cudaMalloc stalls for 50 - 150 ms every so often (whatever its size), which the
            # first run of a model may or may not meet; the second finds its device memory in the library's cache.  Both are
            # reported; ms_e2e is the faster of the two whole runs.
            t1 = time.perf_counter()
            st = m.run(device=device, seed=2)
            t2 = time.perf_counter()
            if first_ms < (t2 - t1) * 1e3:
                t1 = t2 - first_ms * 1e-3
            m.export(path, t2 - t1)            # ss_*.txt / per_*.txt next to the model file
            t3 = time.perf_counter()
            k_ms = st.kernel_ms
            rec = {"model": name, "phonons": int(st.total_phonons), "cells": int(m.info.num_cells), "sensors": int(m.info.num_sensors),
                   "measurement_steps": int(m.info.measurement_steps), "sim_type": int(m.info.sim_type), "load_ms": round(load_ms, 2),
                   "ms_e2e": round((t2 - t1) * 1e3, 2), "ms_e2e_first_run": round(first_ms, 2), "export_ms": round((t3 - t2) * 1e3, 2), "kernel_ms": round(k_ms, 2), "launches": int(st.launches),
                   "drift_steps": int(st.drift_steps), "segments": int(st.events),
                   "drift_steps_per_s": st.drift_steps / (k_ms * 1e-3), "segments_per_s": st.events / (k_ms * 1e-3),
                   "reference_header_s": REFERENCE_HEADER_SECONDS.get(name)}
            out.append(rec)
            m.close()
    return out


def run_ours(args):
    env = Env(args)
    world, rank = env.world, env.rank
    total = args.phonons
    per_gpu = total // world
    model_dict = workload_model(total)
    opts = {"tally_shared": args.tally_shared, "warps_per_sm": args.warps_per_sm if args.warps_per_sm > 0 else -1, "kernel": args.kernel}

    sampler = ClockSampler(env.local) if rank == 0 else None
    job = ShardedJob(env, model_dict, args, opts)
    main = job.timed_jobs(args.warmup, args.steps, 1000, sampler)
    strong = {"si_ge_grid": job.checksum_job()}
    M, S, R, cells = job.M, job.S, job.R, job.cells
    job.close()

    e2e_ms, e2e_detail, h2d, d2h = end_to_end(env, model_dict, args, max(1, min(args.steps, 3)))
    if sampler:
        sampler.__exit__()  # clocks sampled under load from the first timed job to the last end-to-end run

    weak = None
    if world > 1 and not args.no_weak:
        wjob = ShardedJob(env, workload_model(total * world), args, opts)
        w = wjob.timed_jobs(2, max(1, min(args.steps, 3)), 3000)
        wjob.close()
        weak = {"value": w["drift_steps"] / (w["ms_per_step"] * 1e-3), "unit": "drift-steps/s", "ms_per_step": w["ms_per_step"],
                "phonons_per_gpu": total, "phonons_total": total * world}

    kinked = None
    if not args.no_kinked:
        from tests import cases
        kd = cases.kinked_model()
        if kd is not None:
            kjob = ShardedJob(env, kd, args, opts)
            k = kjob.timed_jobs(1, 2, 5000)
            kinked = kjob.checksum_job()
            kinked.update({"workload": "kinked_demo_120_35_spec.json (BASELINE.json configs[1]): 6174 cells, 3108 sensors, 2e7 phonons in total",
                           "value": k["drift_steps"] / (k["ms_per_step"] * 1e-3), "unit": "drift-steps/s", "ms_per_step": k["ms_per_step"],
                           "kernel_ms": k["kernel_ms"], "segments_per_s": k["events"] / (k["ms_per_step"] * 1e-3)})
            kinked.pop("cuts", None)
            kjob.close()
            strong["kinked"] = kinked

    ms_per_step = main["ms_per_step"]
    total_drift = main["drift_steps"]
    value = total_drift / (ms_per_step * 1e-3)
    out = None
    if rank == 0:
        peak_gbs, sm_max, peak_src = measured_peaks()
        clocks = sampler.summary() if sampler else None
        sm_mhz = (clocks or {}).get("sm_mhz") or sm_max
        k_ms = main["kernel_ms"]
        stats = main["stats"]
        launches_per_job = max(1, stats["launches"])
        auto = args.steps_per_launch == 0 and args.kernel < 0 and args.tally_shared < 0 and args.warps_per_sm <= 0
        ncu = ncu_constants(launches_per_job, main["events"] / total_drift, total) if auto else None
        peak_issue = 148 * 4 * sm_mhz * 1e6
        roof = {"bound": "issue", "unit": "warp-inst/s", "peak": peak_issue, "achieved": None, "frac": None, "traffic": None,
                "peak_source": f"148 SMs x 4 schedulers x {sm_mhz:.0f} MHz (SM clock sampled during the timed region)",
                "kernel": {0: "drift_kernel_slots<4>", 1: "drift_kernel_lockstep"}.get(stats["kernel"], "drift_kernel_queues<128, *>"),
                "avg_launch_ms": k_ms / launches_per_job, "launches_per_job": launches_per_job, "kernel_ms_per_job": k_ms,
                "segments_per_s": main["events"] / (ms_per_step * 1e-3), "segments_per_drift_step": main["events"] / total_drift,
                "hbm": {"bound": "hbm", "unit": "GB/s", "peak": peak_gbs, "peak_source": peak_src, "achieved": None, "frac": None,
                        "algorithmic_bytes_per_drift_step": 64,
                        "algorithmic_ratio": main["drift_steps_rank"] * 64 / (k_ms * 1e-3) / 1e9 / peak_gbs,
                        "note": "achieved = DRAM bytes per job measured by ncu (dram__bytes_read + write over the job's launches) / live kernel "
                                "time.  algorithmic_ratio = 64 B x drift-steps / kernel time / peak (SURVEY 8d's accounting) is NOT a roofline "
                                "fraction: a launch keeps a phonon on chip for a whole window of measurement steps"},
                "ncu": ("profiles/r02_ncu_summary.json (device_sha16 matches this tree; launches and segments per drift-step of this job match the captured one)" if ncu else
                        "ncu-derived fields are null: no capture of this source tree / configuration is committed")}
        if ncu:
            j = ncu["bench_job"]
            inst = j["warp_instructions_per_drift_step"] * main["drift_steps_rank"]
            roof["achieved"] = inst / (k_ms * 1e-3)
            roof["frac"] = roof["achieved"] / peak_issue
            roof["warp_instructions_per_drift_step"] = j["warp_instructions_per_drift_step"]
            roof["warp_instructions_per_segment"] = j["warp_instructions_per_drift_step"] * total_drift / main["events"]
            roof["threads_active_per_instruction"] = j["threads_active_per_instruction"]
            dram = j["dram_bytes_per_drift_step"] * main["drift_steps_rank"]
            roof["traffic"] = int(dram / launches_per_job)
            roof["hbm"]["achieved"] = dram / (k_ms * 1e-3) / 1e9
            roof["hbm"]["frac"] = roof["hbm"]["achieved"] / peak_gbs
        out = {
            "metric": METRIC, "value": value, "unit": "drift-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "phonons_total": total, "phonons_per_gpu": per_gpu, "measurement_steps": M, "cells": cells,
                       "sensors": S, "drift_steps_per_job": total_drift, "launch_windows": main["cuts"],
                       "sharding": f"phonon id mod {world}", "l2": "inputs larger than L2 (live pool >> 126 MB, streamed every launch)"
                       if per_gpu >= 50_000_000 else "live pool of a rank below the 126 MB L2 at this N; jobs alternate between two pool copies and "
                       "re-create every phonon, nothing is reused between timed jobs",
                       "rng": "Philox4x32-10 keyed by (seed, phonon id, step)"},
            "roofline": roof,
            "e2e": {"value": total_drift / (e2e_ms * 1e-3), "unit": "drift-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "runs": e2e_detail},
            "gpu_launches": int(main["launches"]),
            "clocks": clocks,
            "strong": strong,
            "weak": weak,
            "stats": stats,
            "csrc_sha16": csrc_sha16(),
            "device_sha16": device_sha16(),
        }
        if world == 1 and not args.no_models:
            out["models"] = model_walltimes(env.local)
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            per_phonon = oracle_drift_steps_per_phonon(model_dict)
            cpu = reference_cpu_run(model_dict, args.cpu_phonons_per_core, cores, per_phonon)
            if cpu and cpu.get("drift_steps_per_s"):
                out["cpu_baseline"] = {"value": cpu["drift_steps_per_s"], "unit": "drift-steps/s", "cores": cores, "kind": cpu["kind"],
                                       "sample": f"{cores} processes x {args.cpu_phonons_per_core} phonons of the same model ({cpu['seconds']:.1f} s); "
                                                 f"{per_phonon:.2f} drift-steps per phonon counted on the CPU by oracle/sim.c on 1e5 phonons"
                                                 if cpu["kind"] == "reference" else f"oracle/sim.c, {cores} threads, counts its own drift-steps"}
            else:
                out["cpu_baseline"] = {"value": None, "unit": "drift-steps/s", "cores": cores, "kind": "reference",
                                       "sample": "neither oracle/_ref/psim_ref nor the oracle library is usable on this box"}
        emit(out)
    if world > 1:
        env.dist.barrier()
        env.dist.destroy_process_group()
    return out


# ---------------------------------------------------------------------------------------------------- reference
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    model = workload_model(args.phonons)
    per_phonon = args.ref_drift_steps_per_phonon if args.ref_drift_steps_per_phonon > 0 else oracle_drift_steps_per_phonon(model)
    vals, secs = [], []
    kind = "reference"
    for it in range(args.warmup + args.steps):
        r = reference_cpu_run(model, args.cpu_phonons_per_core, cores, per_phonon)
        if r is None or not r.get("drift_steps_per_s"):
            emit({"impl": "reference", "unavailable": "reference run failed"})
            return
        kind = r["kind"]
        if it >= args.warmup:
            vals.append(r["drift_steps_per_s"])
            secs.append(r["seconds"])
    v = float(np.mean(vals))
    sample = (f"{cores} processes x {args.cpu_phonons_per_core} phonons of the same model per step; {per_phonon:.2f} drift-steps per "
              f"phonon, counted on the CPU by the restatement of the reference's event loop (oracle/sim.c) on 1e5 phonons of this model"
              if kind == "reference" else f"oracle/sim.c with {cores} threads, counting its own drift-steps")
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "drift-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "phonons_per_step": args.cpu_phonons_per_core * cores},
        "cpu_baseline": {"value": v, "unit": "drift-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "drift-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--phonons", type=int, default=100_000_000, help="phonons of the job IN TOTAL (sharded over the GPUs)")
    ap.add_argument("--steps-per-launch", type=int, default=0, help="0 = library default (automatic)")
    ap.add_argument("--reduce-every", type=int, default=0,
                    help="recorded measurement steps per tally all-reduce group (0 = one group per launch window of the library)")
    ap.add_argument("--tally-shared", type=int, default=-1)
    ap.add_argument("--warps-per-sm", type=int, default=0)
    ap.add_argument("--kernel", type=int, default=-1, help="2 work queues (default), 0 lane-bound slots, 1 lock-step first version")
    ap.add_argument("--cpu-phonons-per-core", type=int, default=400_000)
    ap.add_argument("--ref-drift-steps-per-phonon", type=float, default=0.0, help="0 = count them with oracle/sim.c (default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-models", action="store_true")
    ap.add_argument("--no-weak", action="store_true")
    ap.add_argument("--no-kinked", action="store_true")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
