#!/usr/bin/env python
"""Installs the UNMODIFIED reference (OtsoBear/PyQMD) into baseline/_ref/ so that it travels to the
GPU box with the repo snapshot (baseline/_ref/ is git-ignored, not gpurun-ignored).

    python baseline/install_reference.py [--src /root/reference]

The reference is five loose .py files with no setup.py / pyproject.toml and its directory is
read-only, so the sources are copied to a scratch directory under /tmp, given a three-line
pyproject.toml that lists them as py-modules, and installed with

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref <scratch>

No file of the reference is edited; nothing is copied into the tracked tree.  The modules are used by
`bench.py --impl reference` (times the reference's own update_particles_cpu / should_decay on the host
cores), by cpu_baseline (kind "reference") and by tests/test_gpu_reference_app.py.
"""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
MODULES = ("particles", "decay_chains", "nuclear_forces", "nuclear_sim", "rendering")


def installed():
    return all(os.path.isfile(os.path.join(DEST, m + ".py")) for m in MODULES)


def install(src="/root/reference", force=False):
    if installed() and not force:
        return DEST
    if not os.path.isfile(os.path.join(src, "nuclear_forces.py")):
        raise RuntimeError(f"reference sources not found under {src}")
    tmp = tempfile.mkdtemp(prefix="pyqmd_ref_")
    try:
        for m in MODULES:
            shutil.copy(os.path.join(src, m + ".py"), tmp)
        with open(os.path.join(tmp, "pyproject.toml"), "w") as f:
            f.write('[build-system]\nrequires = ["setuptools"]\nbuild-backend = "setuptools.build_meta"\n'
                    '[project]\nname = "pyqmd-reference"\nversion = "0"\n'
                    "[tool.setuptools]\npy-modules = [%s]\n" % ", ".join(f'"{m}"' for m in MODULES))
        if os.path.isdir(DEST):
            shutil.rmtree(DEST)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", DEST, tmp]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n" + res.stdout + res.stderr)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    for m in MODULES:       # the installed modules must be byte-identical to the reference's
        with open(os.path.join(src, m + ".py"), "rb") as a, open(os.path.join(DEST, m + ".py"), "rb") as b:
            if a.read() != b.read():
                raise RuntimeError(f"{m}.py differs from the reference after installation")
    return DEST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    print(install(a.src, a.force))
