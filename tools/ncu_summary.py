#!/usr/bin/env python
"""profiles/<tag>_ncu_summary.json (what bench.py reads for roofline.traffic) and the raw-page CSVs, from the captures
tools/gpu_evidence.sh leaves in gpurun_out/evidence.  usage: python tools/ncu_summary.py [tag]   (here, no GPU)"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EV = os.path.join(ROOT, "gpurun_out", "evidence")
PROFILES = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

CAPTURES = {
    "long": ("queues_long_window", "bench workload, 899-step window (first launch of a job: nothing is recorded before step 900)"),
    "rec": ("queues_recorded_window", "bench workload, 36-step recorded window (tallies staged in shared memory, difference form)"),
    "per": ("queues_periodic_global", "linear_sides periodic, 1000 sensors: 128-step recorded window, tallies straight to global memory (difference form)"),
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
    "smsp__inst_executed_op_global_red.sum": "global_red_instructions",
    "lts__t_sectors_srcunit_tex_op_red.sum": "l2_red_sectors",
    "lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed": "l2_red_pct_of_peak",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard_per_issue",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait_per_issue",
}
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}


def raw_page(name):
    rep = os.path.join(EV, f"prof_{name}.ncu-rep")
    if not os.path.exists(rep):
        return None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    with open(os.path.join(PROFILES, f"{tag}_ncu_{CAPTURES[name][0]}.csv"), "w") as f:
        f.write(out)
    return dict(zip(rows[0], zip(rows[2], rows[1])))


def main():
    captures = []
    for name, (_, what) in CAPTURES.items():
        d = raw_page(name)
        if d is None:
            continue
        c = {"what": what, "kernel": d["Kernel Name"][0].split("(")[0].replace("void <unnamed>::", "")}
        for metric, key in KEEP.items():
            if metric in d:
                v, unit = d[metric]
                x = float(v)
                if key.startswith("dram_bytes"):
                    x *= SCALE.get(unit, 1.0)
                    x = int(x)
                elif key in ("warp_instructions", "shared_atomic_instructions", "shared_atomic_wavefronts", "global_red_instructions",
                             "l2_red_sectors", "registers_per_thread"):
                    x = int(x)
                c[key] = x
        captures.append(c)
    # DRAM bytes of every launch of one job of the bench workload
    launches = {}
    path = os.path.join(EV, f"{tag}_dram_per_launch.csv")
    for row in csv.reader(l for l in open(path) if l.startswith('"')):
        if row[0] == "ID":
            continue
        launches.setdefault(int(row[0]), {})[row[12]] = float(row[14])
    per_launch = [{"launch": i, "dram_bytes": int(v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]),
                   "ms_under_ncu": v["gpu__time_duration.sum"] * 1e-6} for i, v in sorted(launches.items())]
    recorded = per_launch[1:]
    summary = {
        "round": int(tag[1:]), "kernel": "drift_kernel_queues<128>",
        "grid": "148 CTAs x 768 threads (one CTA per SM), 80 registers/thread",
        "commands": {
            "long / rec": "ncu --set full --clock-control none --import-source on -k regex:drift_kernel -s 0|1 -c 1 python bench.py --steps 1 --warmup 0 --no-cpu-baseline",
            "per": "same with -s 2 on python tools/profile_model.py sides_per",
            "dram_per_launch": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:drift_kernel -c 4 ({tag}_dram_per_launch.csv)",
        },
        "phonons_per_gpu": 100_000_000,
        "captures": captures,
        "default_job": {
            "launches": len(per_launch), "per_launch": per_launch,
            "dram_bytes_long_window": per_launch[0]["dram_bytes"],
            "dram_bytes_recorded_window": int(sum(p["dram_bytes"] for p in recorded) / max(len(recorded), 1)),
            "dram_bytes_per_launch_avg": int(sum(p["dram_bytes"] for p in per_launch) / len(per_launch)),
        },
    }
    with open(os.path.join(PROFILES, f"{tag}_ncu_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
