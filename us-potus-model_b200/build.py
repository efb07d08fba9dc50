"""nvcc recipe for the in-tree CUDA library (sm_100a only; no other arch, no fallback path)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libpotus_b200.so")
SOURCES = ["potus_lib.cu", "potus_kernel.cu", "potus_stream.cu", "potus_post.cu", "potus_host.cu", "potus_stream_host.cuh", "potus_layout.h",
           "potus_stream_layout.h", "ptx_sm100.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=hidden", "--shared", "-Xptxas", "-v"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the CUDA library cannot be built (there is no CPU fallback)")
    return exe


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(os.path.dirname(HERE), "include", "potus_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_prof() -> str:
    """Development build with per-phase clock64 counters (-DPOTUS_PROF); never used by tests or bench."""
    os.makedirs(LIB_DIR, exist_ok=True)
    out = os.path.join(LIB_DIR, "libpotus_b200_prof.so")
    res = subprocess.run([_nvcc(), *NVCC_FLAGS, "-DPOTUS_PROF", "-o", out, os.path.join(CSRC, "potus_lib.cu")], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    tmp = LIB + f".tmp{os.getpid()}"          # build aside, then rename: a concurrent reader never sees a half-written library
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", tmp, os.path.join(CSRC, "potus_lib.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB)
    cmd[cmd.index(tmp)] = LIB
    log = os.path.join(LIB_DIR, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB
