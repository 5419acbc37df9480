"""Builds libpyqmd_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m pyqmd_b200.build [--force] [--verbose]

The shared library is a plain C-ABI object (include/pyqmd_b200.h); it is git-ignored but
travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libpyqmd_b200.so")
SOURCES = ["api.cu", "cloud.cu", "cloud_host.cu", "cloud_sym.cu", "ensemble.cu", "free_particles.cu", "layout.cu", "overlaps.cu", "population.cu"]
HEADERS = ["common.cuh", "pair_law.cuh", "cloud.cuh", "decay_device.cuh", "../../include/pyqmd_b200.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--extra-device-vectorization",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libpyqmd_b200.so cannot be built")


def is_stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = SO) -> str:
    """``defines`` / ``out``: tuning builds (A/B variants selected with PYQMD_B200_LIB)."""
    if not force and not is_stale() and out == SO:
        return SO
    cmd = [nvcc_path(), *NVCC_FLAGS, *[f"-D{d}" for d in defines]]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, f) for f in SOURCES] + ["-o", out]
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    res = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libpyqmd_b200.so")
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv or bool(defs), verbose="--verbose" in sys.argv,
                defines=defs, out=os.path.abspath(outs[0]) if outs else SO))
