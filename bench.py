#!/usr/bin/env python
"""Benchmark of the phonon Monte Carlo particle loop (BASELINE.json: phonon drift-steps/sec + HBM roofline %).

Workload (config.workload): BASELINE.json configs[4], the synthetic 100-cell Si/Ge structure with 1e8 deviational
phonons PER GPU (weak scaling: an N-GPU job simulates N x 1e8 phonons, rank r owns the phonon ids == r mod N),
1000 measurement steps over 1 ns.  One bench "step" = one complete simulation of that model: every measurement
interval of every phonon (emission, drift, scattering, surfaces, cell transitions, tallies) plus the all-reduce of
the sensor tallies.  1 drift-step = one live phonon advanced across one measurement interval.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference ...                         the reference's own CPU code on the host cores

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream, barrier + synchronize on both sides,
max over ranks.  The phonon pool (~0.9 GB live per launch) is far larger than the 126 MB L2, so every launch
streams it from HBM (config.l2: "inputs larger than L2").
"""
from __future__ import annotations

import argparse
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

ALGO_BYTES_PER_DRIFT_STEP = 64  # 32 B state read + 32 B written
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "psim_ref")
METRIC = "phonon drift-steps/sec"


def workload_model(num_phonons: int) -> dict:
    return configs.si_ge_grid(num_phonons=num_phonons).to_dict()


def ncu_traffic_per_launch(per_gpu: int, launches_per_job: int, auto_windows: bool):
    """DRAM bytes per drift-kernel launch, averaged over the launches of one job, from the committed ncu captures
    (profiles/r01_ncu_summary.json: one long unrecorded window + recorded windows); None if they were taken on another
    configuration."""
    path = os.path.join(ROOT, "profiles", "r01_ncu_summary.json")
    try:
        d = json.load(open(path))
        if d["phonons_per_gpu"] == per_gpu and auto_windows and launches_per_job >= 2:
            j = d["default_job"]
            return int((j["dram_bytes_long_window"] + (launches_per_job - 1) * j["dram_bytes_recorded_window"]) / launches_per_job)
    except Exception:
        pass
    return None


def issue_slot_use(per_gpu: int, kernel_ms: float, recorded_steps: int, sm_mhz, auto_windows: bool):
    """What actually bounds the kernel: warp instructions issued per second against the issue
    slots of the chip (148 SMs x 4 schedulers x SM clock).  Instruction counts per launch come from the committed ncu
    captures of this workload (one long unrecorded window + one 36-step recorded window, scaled to the recorded steps of
    a job); the time is the one measured live.  None when the captures do not describe this configuration."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_summary.json")))
        if d["phonons_per_gpu"] != per_gpu or not auto_windows or not sm_mhz:
            return None
        long_w, rec_w = d["captures"][0], d["captures"][1]
        inst = long_w["warp_instructions"] + rec_w["warp_instructions"] * recorded_steps / 36.0
        peak = 148 * 4 * float(sm_mhz) * 1e6
        achieved = inst / (kernel_ms * 1e-3)
        return {"achieved_warp_inst_per_s": achieved, "peak_warp_inst_per_s": peak, "frac": achieved / peak,
                "threads_active_per_instruction": long_w["threads_active_per_instruction"],
                "source": "warp instructions per launch from profiles/r01_ncu_summary.json (ncu), time measured live"}
    except Exception:
        return None


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


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
def reference_cpu_run(model: dict, phonons_per_proc: int, procs: int, drift_steps_per_phonon: float | None):
    """Runs the reference's own CPU implementation (oracle/_ref/psim_ref: unmodified reference sources + our
    driver main) as `procs` independent processes (no TBB in this image, so std::execution::par is serial; phonons
    are independent, so P processes of n/P phonons are the reference's parallel path).  Wall = slowest process."""
    if not os.path.exists(REF_BIN):
        return oracle_cpu_run(model, phonons_per_proc * procs, procs, drift_steps_per_phonon)
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


def oracle_cpu_run(model: dict, phonons: int, threads: int, drift_steps_per_phonon):
    """Fallback when the reference binary is not on the box: the plain-C restatement under oracle/ (OpenMP)."""
    try:
        from oracle.model import OracleModel
        om = OracleModel(configs.with_settings(model, num_phonons=phonons))
        om.prepare()
        t0 = time.perf_counter()
        _, _, steps, _ = om.run(1, threads=threads)
        secs = time.perf_counter() - t0
    except Exception:
        return None
    return {"phonons": phonons, "seconds": secs, "wall_with_load": secs, "kind": "port",
            "drift_steps_per_s": (phonons * drift_steps_per_phonon if drift_steps_per_phonon else steps) / secs}


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


def run_ours(args):
    import torch
    import torch.distributed as dist

    from psim_b200 import lib as psim

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    per_gpu = args.phonons
    model_dict = workload_model(per_gpu * world)
    model = psim.Model(text=json.dumps(model_dict))
    model.prepare()
    info = model.info
    M, S, R = info.measurement_steps, info.num_sensors, info.recorded_steps
    desc = model.describe()

    g = psim.GpuSimulator(desc, local)
    g.set_option("steps_per_launch", args.steps_per_launch)
    if args.tally_shared >= 0:
        g.set_option("tally_shared", args.tally_shared)
    if args.warps_per_sm > 0:
        g.set_option("warps_per_sm", args.warps_per_sm)
    if args.kernel >= 0:
        g.set_option("kernel", args.kernel)

    class _Dev:  # expose the library's tally buffers to torch (for the NCCL all-reduce) without a copy
        def __init__(self, ptr, shape, typestr):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}

    e_ptr, f_ptr, _, _ = g.tally_buffers()
    t_energy = torch.as_tensor(_Dev(e_ptr, (R, S), "<i4"), device=f"cuda:{local}")
    t_flux = torch.as_tensor(_Dev(f_ptr, (R, S, 2), "<i8"), device=f"cuda:{local}")
    stream = torch.cuda.Stream(device=local)  # the drift kernels are launched on THIS stream, and so are the events
    torch.cuda.set_stream(stream)
    chunk = max(args.steps_per_launch, args.reduce_every) if args.reduce_every > 0 else 0

    first = M - R  # first step whose measurement is recorded

    def plan_cuts():
        """Steps whose measurement is not recorded (steady state: the first 90 %) need no exchange and go to the library
        in one call (it chooses its own launch windows); the recorded steps go in groups that end where the library's
        launch windows end (no extra launch for the exchange; --reduce-every N asks for groups of N steps instead)."""
        cuts, s = [0], 0
        if first > 1:
            s = first - 1
            cuts.append(s)
        while s < M - 1:
            s = min(s + chunk, M - 1) if chunk > 0 else g.next_window(s)
            cuts.append(s)
        return agree_on_cuts(cuts, M, world, rank, f"cuda:{local}")

    def one_job(cuts):
        """All measurement steps; the tally all-reduce of a group of steps is issued as soon as the group's launches are
        enqueued."""
        for s, e in zip(cuts[:-1], cuts[1:]):
            g.run_steps(s, e, stream.cuda_stream)
            if world > 1:
                r0, r1 = max(s + 1 - first, 0), e + 1 - first  # tally rows completed by steps [s, e)
                if r1 > r0:
                    dist.all_reduce(t_energy[r0:r1])
                    dist.all_reduce(t_flux[r0:r1])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    times, kernel_ms, drift, launches = [], [], [], 0
    sampler = ClockSampler(local) if rank == 0 else None
    total_steps = args.warmup + args.steps
    for it in range(total_steps):
        seed = 1000 + it
        src, n = model.sources(seed)
        g.set_sources(src, n, seed, rank, world)  # untimed: pool reset, tallies zeroed, birth plan resident in HBM
        cuts = plan_cuts()
        timed = it >= args.warmup
        if timed and it == args.warmup and sampler:
            sampler.__enter__()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        one_job(cuts)
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        st = g.stats()
        if timed:
            times.append(ms)
            kernel_ms.append(st.kernel_ms)
            drift.append(st.drift_steps)
            launches += st.launches
    last_stats = g.stats().as_dict()

    # ---- end to end through the host API with HOST buffers, as the runs of a multi-run model go (psim_model_run):
    # per run the sources / birth plan go host -> device, the GPU runs, the tallies come back device -> host and the
    # reference's run epilogue (temperatures / fluxes) is done on the host.  The handle (model image, pool) is created
    # once, as psim_model_run does; the first run through it is the warm-up. ----
    e2e_ms, e2e_detail = [], []
    h2d = d2h = 0
    g2 = psim.GpuSimulator(model.describe(), local)  # uploads cells / sensors / tables (host -> device), once
    g2.set_option("steps_per_launch", args.steps_per_launch)
    for it in range(1 + max(1, min(args.steps, 3))):
        seed = 2000 + it
        barrier()
        t0 = time.perf_counter()
        src, n = model.sources(seed)
        g2.set_sources(src, n, seed, rank, world)  # birth plan host -> device
        g2.run()
        e, f = g2.tallies()  # device -> host
        if world > 1:
            te, tf = torch.from_numpy(e.astype(np.int64)).cuda(), torch.from_numpy(f).cuda()
            dist.all_reduce(te)
            dist.all_reduce(tf)
            e, f = te.cpu().numpy().astype(np.int32), tf.cpu().numpy()
        model.set_tallies(e, f)
        model.finish_run(0)
        six, _, _ = model.results(0, traces=False)
        model.next_run()
        model.prepare()
        barrier()
        ms = (time.perf_counter() - t0) * 1e3
        st2 = g2.stats()
        e2e_detail.append({"ms": round(ms, 2), "kernel_ms": round(st2.kernel_ms, 2), "phonons": st2.total_phonons,
                           "drift_steps": st2.drift_steps, "timed": it > 0})
        if it > 0:
            e2e_ms.append(ms)
        h2d = st2.plan_bytes  # sources + birth plan, counted by the library
        d2h = st2.tally_bytes
    g2.close()
    if sampler:
        sampler.__exit__()  # clocks sampled under load from the first timed job to the last end-to-end run

    t_max = torch.tensor([float(np.mean(times)), float(np.mean(e2e_ms))], device=f"cuda:{local}")
    d_sum = torch.tensor([float(np.mean(drift))], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(d_sum, op=dist.ReduceOp.SUM)
    ms_per_step = float(t_max[0])
    e2e_ms_max = float(t_max[1])
    total_drift = float(d_sum[0])
    value = total_drift / (ms_per_step * 1e-3)

    out = None
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        k_ms = float(np.mean(kernel_ms))
        achieved = float(np.mean(drift)) * ALGO_BYTES_PER_DRIFT_STEP / (k_ms * 1e-3) / 1e9
        launches_per_job = max(1, last_stats["launches"])
        out = {
            "metric": METRIC, "value": value, "unit": "drift-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "synthetic 100-cell Si/Ge 2D structure (BASELINE.json configs[4]), steady-state deviational",
                       "phonons_per_gpu": per_gpu, "phonons_total": per_gpu * world, "measurement_steps": M, "cells": info.num_cells,
                       "sensors": S, "drift_steps_per_job": total_drift, "steps_per_launch": last_stats["steps_per_launch"],
                       "sharding": f"phonon id mod {world}", "l2": "inputs larger than L2 (live pool >> 126 MB, streamed every launch)",
                       "rng": "Philox4x32-10 keyed by (seed, phonon id, step)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_per_launch(per_gpu, launches_per_job, args.steps_per_launch == 0),
                         "peak_source": peak_src, "kernel": {0: "drift_kernel_slots<4>", 1: "drift_kernel_lockstep"}.get(last_stats["kernel"], "drift_kernel_queues<128>"),
                         "algorithmic_bytes_per_drift_step": ALGO_BYTES_PER_DRIFT_STEP,
                         "algorithmic_bytes_per_launch": float(np.mean(drift)) * ALGO_BYTES_PER_DRIFT_STEP / launches_per_job,
                         "avg_launch_ms": k_ms / launches_per_job, "launches_per_job": launches_per_job, "kernel_ms_per_job": k_ms,
                         "issue_slots": issue_slot_use(per_gpu, k_ms, R, (sampler.summary() or {}).get("sm_mhz") if sampler else None,
                                                       args.steps_per_launch == 0),
                         "note": "achieved = 64 B x drift-steps / kernel time, the accounting of SURVEY 8d (state read + written once per "
                                 "drift-step).  A launch keeps a phonon on chip for a whole window of measurement steps, so the real "
                                 "DRAM traffic (`traffic`, bytes per launch, ncu) is a small fraction of the algorithmic bytes and frac "
                                 "can exceed 1: the kernel is issue-bound (issue active 71 %), not HBM-bound"},
            "e2e": {"value": total_drift / (e2e_ms_max * 1e-3), "unit": "drift-steps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms_max, "runs": e2e_detail},
            "gpu_launches": int(launches),
            "clocks": sampler.summary() if sampler else None,
            "stats": last_stats,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            per_phonon = float(np.mean(drift)) / per_gpu
            cpu = reference_cpu_run(model_dict, args.cpu_phonons_per_core, cores, per_phonon)
            if cpu:
                out["cpu_baseline"] = {"value": cpu["drift_steps_per_s"], "unit": "drift-steps/s", "cores": cores,
                                       "kind": cpu["kind"],
                                       "sample": f"{cores} processes x {args.cpu_phonons_per_core} phonons of the same model "
                                                 f"({cpu['seconds']:.1f} s); drift-steps per phonon taken from the GPU run"}
            else:
                out["cpu_baseline"] = {"value": None, "unit": "drift-steps/s", "cores": cores, "kind": "reference",
                                       "sample": "oracle/_ref/psim_ref not present on this box"}
        emit(out)
    g.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


# ---------------------------------------------------------------------------------------------------- reference
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    model = workload_model(args.phonons * args.gpus)
    per_phonon = args.ref_drift_steps_per_phonon
    vals, secs = [], []
    kind = "reference"
    for it in range(args.warmup + args.steps):
        r = reference_cpu_run(model, args.cpu_phonons_per_core, cores, per_phonon)
        if r is None:
            emit({"impl": "reference", "unavailable": "reference run failed"})
            return
        kind = r["kind"]
        if it >= args.warmup:
            vals.append(r["drift_steps_per_s"])
            secs.append(r["seconds"])
    v = float(np.mean(vals))
    sample = (f"{cores} processes x {args.cpu_phonons_per_core} phonons of the same model per step; "
              f"{per_phonon} drift-steps per phonon (counted by the CUDA path on this model)")
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "drift-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic 100-cell Si/Ge 2D structure (BASELINE.json configs[4]), steady-state deviational",
                   "phonons_per_step": args.cpu_phonons_per_core * cores},
        "cpu_baseline": {"value": v, "unit": "drift-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "drift-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--phonons", type=int, default=100_000_000, help="phonons per GPU")
    ap.add_argument("--steps-per-launch", type=int, default=0, help="0 = library default (automatic)")
    ap.add_argument("--reduce-every", type=int, default=0,
                    help="recorded measurement steps per tally all-reduce group (0 = one group per launch window of the library)")
    ap.add_argument("--tally-shared", type=int, default=-1)
    ap.add_argument("--warps-per-sm", type=int, default=0)
    ap.add_argument("--kernel", type=int, default=-1, help="2 work queues (default), 0 lane-bound slots, 1 lock-step first version")
    ap.add_argument("--cpu-phonons-per-core", type=int, default=400_000)
    ap.add_argument("--ref-drift-steps-per-phonon", type=float, default=133.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
