"""Build recipe: nvcc -> psim_b200/lib/libpsim_b200.so (sm_100a kernels + C ABI + host layer) and the `psim` CLI.

In-tree on purpose: the built library travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
BIN_DIR = os.path.join(HERE, "bin")
LIB = os.path.join(LIB_DIR, "libpsim_b200.so")
CLI = os.path.join(BIN_DIR, "psim")
EMU = os.path.join(ROOT, "tests", "emu", "libpsim_emu.so")

SOURCES = ["psim_gpu.cu", "flatten.cpp", "host/model.cpp", "host/host_api.cpp"]
HEADERS = ["device_core.cuh", "kernels.cuh", "device_types.h", "flatten.h", "host/model.h", "host/json.h",
           "../../include/psim_b200.h", "../../include/psim_host.h"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(BIN_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    if force or _stale(LIB, deps):
        cmd = [_nvcc(), *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-Wall", "-shared",
               "-o", LIB, *srcs, "-ldl"]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        subprocess.run(cmd, check=True)
    main_src = os.path.join(CSRC, "host", "main.cpp")
    if force or _stale(CLI, [main_src, LIB]):
        subprocess.run(["g++", "-O2", "-std=c++17", "-o", CLI, main_src, "-L" + LIB_DIR, "-lpsim_b200",
                        "-Wl,-rpath,$ORIGIN/../lib"], check=True)
    return LIB


def build_variant(name: str, defines: dict) -> str:
    """Tuning helper: the same library with -D overrides (e.g. scheduler thresholds) under lib/variants/."""
    out_dir = os.path.join(LIB_DIR, "variants")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"libpsim_b200_{name}.so")
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    cmd = [_nvcc(), *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-o", out,
           *[f"-D{k}={v}" for k, v in defines.items()], *srcs, "-ldl"]
    subprocess.run(cmd, check=True)
    return out


def build_emu(force: bool = False) -> str:
    """Test-only: the device core compiled for the host (tests/emu). Never loaded by the package."""
    src = os.path.join(ROOT, "tests", "emu", "emu.cpp")
    deps = [src, os.path.join(CSRC, "device_core.cuh"), os.path.join(CSRC, "device_types.h"),
            os.path.join(CSRC, "flatten.cpp"), os.path.join(CSRC, "flatten.h")]
    if force or _stale(EMU, deps):
        cuda_inc = os.path.join(os.path.dirname(os.path.dirname(_nvcc())), "include")
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
                        "-I" + cuda_inc, "-o", EMU, src, os.path.join(CSRC, "flatten.cpp")], check=True)
    return EMU


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
