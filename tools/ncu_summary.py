#!/usr/bin/env python
"""profiles/<tag>_ncu_summary.json (what bench.py reads for its roofline block - only while `device_sha16` matches the tree that
runs), the raw-page CSVs of the --set full captures and the SASS opcode histogram of the dominant kernel, from what
tools/gpu_evidence.sh left in gpurun_out/evidence.  usage: python tools/ncu_summary.py [tag]   (here, no GPU)"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EV = os.path.join(ROOT, "gpurun_out", "evidence")
PROFILES = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"

CAPTURES = {
    "long": ("queues_long_window", "bench workload (Si/Ge, 1e8 phonons), 899-step window over the lattice image: first launch of a job, nothing is recorded before step 900"),
    "rec": ("queues_recorded_window", "bench workload, 36-step recorded window (tallies staged in shared memory, difference form)"),
    "per": ("queues_periodic_global", "linear_sides periodic, 1000 sensors: 128-step recorded window, tallies posted to global memory sector by sector"),
    "kinked": ("queues_kinked_long_window", "kinked wire (6174 cells -> 3150 flight cells -> 250 lattice cells), 1023-step unrecorded window over the lattice image"),
    "kinked_rec": ("queues_kinked_recorded_window", "kinked wire, 128-step recorded window over the lattice image (3108 sensors: every crossed measurement attributed to its sensor area, posted to global memory)"),
}
KEEP = {
    "gpu__time_duration.sum": "launch_ms_under_ncu",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_active_per_instruction",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_kb",
    "launch__registers_per_thread": "registers_per_thread",
    "smsp__inst_executed_op_shared_atom.sum": "shared_atomic_instructions",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum": "shared_atomic_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflict_wavefronts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_wavefronts",
    "smsp__inst_executed_op_global_red.sum": "global_red_instructions",
    "lts__t_sectors_srcunit_tex_op_red.sum": "l2_red_sectors",
    "lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed": "l2_red_pct_of_peak",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard_per_issue",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait_per_issue",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction_per_issue",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected_per_issue",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch_resolving_per_issue",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle_per_issue",
}
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}


def number(text):
    try:
        return float(text.replace(",", ""))
    except ValueError:
        return None


def raw_page(name):
    rep = os.path.join(EV, f"prof_{name}.ncu-rep")
    if not os.path.exists(rep):
        return None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    with open(os.path.join(PROFILES, f"{tag}_ncu_{CAPTURES[name][0]}.csv"), "w") as f:
        f.write(out)
    return dict(zip(rows[0], zip(rows[2], rows[1])))


def sass_histogram(name, path):
    rep = os.path.join(EV, f"prof_{name}.ncu-rep")
    if not os.path.exists(rep):
        return
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    ix = {h: i for i, h in enumerate(rows[1])}
    insts = rows[2:]
    tot = sum(int(r[ix["Instructions Executed"]]) for r in insts)
    by = collections.defaultdict(lambda: [0, 0, 0])
    for r in insts:
        parts = r[ix["Source"]].split()
        op = (parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "?")).split(".")[0]
        by[op][0] += 1
        by[op][1] += int(r[ix["Instructions Executed"]])
        by[op][2] += int(r[ix["Thread Instructions Executed"]])
    with open(path, "w") as f:
        f.write(f"# SASS opcode histogram of {rows[0][1]}\n# capture: {CAPTURES[name][1]}\n")
        f.write(f"# static instructions {len(insts)}, executed warp instructions {tot}\n")
        f.write(f"# {'opcode':12s} {'static':>7s} {'executed':>15s} {'share':>7s} {'avg threads':>12s}\n")
        for op, (n, e, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
            f.write(f"  {op:12s} {n:7d} {e:15d} {100 * e / max(tot, 1):6.2f}% {t / max(e, 1):12.1f}\n")


def job_counters():
    """The four launches of one bench job (tools/gpu_evidence.sh): totals per job and per drift-step."""
    path = os.path.join(EV, f"{tag}_job_counters.csv")
    line = os.path.join(EV, f"{tag}_job_counters_bench.json")
    if not (os.path.exists(path) and os.path.exists(line)):
        return None
    text = [ln for ln in open(path) if not ln.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(text))))
    launches = collections.OrderedDict()
    for r in rows:
        d = launches.setdefault(r["ID"], {"kernel": r["Kernel Name"]})
        v = number(r["Metric Value"])
        d[r["Metric Name"]] = v * SCALE.get(r["Metric Unit"], 1.0) if v is not None else None
    bench = json.loads([ln for ln in open(line) if ln.strip().startswith("{")][-1])
    steps = bench["config"]["drift_steps_per_job"]
    ls = list(launches.values())
    inst = sum(x["smsp__inst_executed.sum"] for x in ls)
    thr = sum(x["smsp__thread_inst_executed.sum"] for x in ls)
    dram = sum(x["dram__bytes_read.sum"] + x["dram__bytes_write.sum"] for x in ls)
    open(os.path.join(PROFILES, f"{tag}_job_counters.csv"), "w").write("".join(text))
    return {"what": "the launches of ONE bench job (Si/Ge, 1e8 phonons, automatic windows) under ncu: per-launch counters in "
                    f"profiles/{tag}_job_counters.csv; the job's drift-steps are counted by the run itself",
            "launches": len(ls), "drift_steps_per_job": steps, "segments_per_job": bench["roofline"]["segments_per_drift_step"] * steps,
            "warp_instructions_per_job": inst, "warp_instructions_per_drift_step": inst / steps,
            "warp_instructions_per_segment": inst / (bench["roofline"]["segments_per_drift_step"] * steps),
            "threads_active_per_instruction": thr / inst, "dram_bytes_per_job": dram, "dram_bytes_per_drift_step": dram / steps,
            "per_launch": [{"kernel": x["kernel"][:60], "ms_under_ncu": x["gpu__time_duration.sum"], "warp_instructions": x["smsp__inst_executed.sum"],
                            "dram_bytes": x["dram__bytes_read.sum"] + x["dram__bytes_write.sum"]} for x in ls]}


def main():
    sha = open(os.path.join(EV, f"{tag}_csrc_sha16.txt")).read().strip()
    dev_sha = open(os.path.join(EV, f"{tag}_device_sha16.txt")).read().strip()
    captures = []
    for name, (_, what) in CAPTURES.items():
        d = raw_page(name)
        if d is None:
            continue
        c = {"capture": name, "what": what}
        for metric, key in KEEP.items():
            if metric in d:
                v, unit = d[metric]
                x = number(v)
                c[key] = x * SCALE.get(unit, 1.0) if (x is not None and key.startswith("dram_bytes")) else x
        captures.append(c)
    out = {"tag": tag, "csrc_sha16": sha, "device_sha16": dev_sha, "phonons_per_gpu": 100_000_000,
           "note": "bench.py uses bench_job only while device_sha16 equals the hash of the device sources (bench.py: DEVICE_SOURCES) of the "
                   "tree that is running and the running job has bench_job's launches and segments per drift-step",
           "bench_job": job_counters(), "captures": captures}
    with open(os.path.join(PROFILES, f"{tag}_ncu_summary.json"), "w") as f:
        json.dump(out, f, indent=1)
    sass_histogram("long", os.path.join(PROFILES, f"{tag}_sass_histogram.txt"))
    for extra in (f"{tag}_launches.csv", f"{tag}_memcheck.log", *[f"{tag}_launches_{m}.csv" for m in ("sige", "kinked", "sides_ss", "sides_per", "linear")]):
        src = os.path.join(EV, extra)
        if os.path.exists(src):
            open(os.path.join(PROFILES, extra), "w").write(open(src).read())
    print(json.dumps({k: v for k, v in out.items() if k != "captures"}, indent=1)[:1500])
    for c in captures:
        print(c["capture"], {k: c.get(k) for k in ("launch_ms_under_ncu", "warp_instructions", "threads_active_per_instruction", "issue_active_pct", "l1_hit_pct")})


if __name__ == "__main__":
    main()
