"""Build recipe: nvcc -> psim_b200/lib/libpsim_b200.so (sm_100a kernels + C ABI + host layer) and the `psim` CLI.

In-tree on purpose: the built library travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
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


def _digest(deps, cmd) -> str:
    # the command line without where the tree or the toolkit happen to live (the GPU box runs a copy under another path)
    def place_free(word: str) -> str:
        flag = word[:2] if word[:2] in ("-I", "-L") else ""
        path = word[len(flag):]
        if path.startswith(ROOT + os.sep):
            return flag + os.path.relpath(path, ROOT)
        return flag + (os.path.basename(path) if os.path.isabs(path) else path)

    words = [place_free(c) for c in cmd]
    h = hashlib.sha256(" ".join(words).encode())
    for d in deps:
        h.update(os.path.basename(d).encode())
        h.update(open(d, "rb").read() if os.path.exists(d) else b"<absent>")
    return h.hexdigest()


def _stale(target: str, deps, cmd) -> bool:
    """By CONTENT, not by mtime: <target>.srchash holds the hash of the inputs and of the command line the target was
    built from, so a checkout, a copy or a touched file can neither hide a change nor force a rebuild."""
    try:
        return not os.path.exists(target) or open(target + ".srchash").read().strip() != _digest(deps, cmd)
    except OSError:
        return True


def _run(target: str, deps, cmd) -> None:
    subprocess.run(cmd, check=True)
    with open(target + ".srchash", "w") as f:
        f.write(_digest(deps, [c for c in cmd if c not in ("-Xptxas", "-v")]) + "\n")


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(BIN_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    cmd = [_nvcc(), *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-Wall", "-shared",
           "-o", LIB, *srcs, "-ldl"]
    if force or _stale(LIB, deps, cmd):
        _run(LIB, deps, cmd if not verbose else cmd[:1] + ["-Xptxas", "-v"] + cmd[1:])
    main_src = os.path.join(CSRC, "host", "main.cpp")
    cli_deps = [main_src, os.path.join(ROOT, "include", "psim_host.h"), LIB + ".srchash"]
    cli_cmd = ["g++", "-O2", "-std=c++17", "-o", CLI, main_src, "-L" + LIB_DIR, "-lpsim_b200",
               "-Wl,-rpath,$ORIGIN/../lib"]
    if force or _stale(CLI, cli_deps, cli_cmd):
        _run(CLI, cli_deps, cli_cmd)
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
    cuda_inc = os.path.join(os.path.dirname(os.path.dirname(_nvcc())), "include")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
           "-I" + cuda_inc, "-o", EMU, src, os.path.join(CSRC, "flatten.cpp")]
    if force or _stale(EMU, deps, cmd):
        _run(EMU, deps, cmd)
    return EMU


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
