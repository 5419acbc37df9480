"""Loader that imports the UNMODIFIED reference (OtsoBear/PyQMD) headlessly.

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/gen_golden.py`` (run in the build
container, where /root/reference exists) to produce golden vectors, and by the
"-m 'not gpu'" tests *when the reference happens to be present* to re-pin the oracle
live; on the GPU box (where only baseline/_ref exists) by tests/test_gpu_reference_app.py and by the
reference arm / cpu_baseline leg of bench.py.  Nothing on the product path imports this.

The reference's ``nuclear_forces.py`` imports ``pyopencl`` at module top
(nuclear_forces.py:2-3) and ``nuclear_sim.py`` imports ``pygame`` and shells out to pip for
``siphash24`` (nuclear_sim.py:1,21-29); none are installed here, so empty stub modules are
injected into ``sys.modules`` first.  ``NuclearForces`` is then built with
``object.__new__`` (bypassing ``setup_opencl``, nuclear_forces.py:19-54) and given the five
strength attributes the constructor would set (nuclear_forces.py:13-17), after which
``update_particles_cpu`` (nuclear_forces.py:236-323) runs exactly as written.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

# baseline/_ref: the unmodified reference installed by baseline/install_reference.py (git-ignored, it
# travels to the GPU box with the snapshot); /root/reference exists in the build container only.
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SEARCH = [os.environ.get("PYQMD_REF", ""), os.path.join(_ROOT, "baseline", "_ref"), "/root/reference"]


def reference_dir():
    if os.environ.get("PYQMD_NO_REF"):          # tests: behave as if no reference were installed
        return None
    for d in _SEARCH:
        if d and os.path.isfile(os.path.join(d, "nuclear_forces.py")):
            return d
    return None


def available() -> bool:
    return reference_dir() is not None


class Ref:
    """Namespace holding the imported reference modules."""

    def __init__(self):
        d = reference_dir()
        if d is None:
            raise RuntimeError("reference sources not found (set PYQMD_REF)")
        sys.dont_write_bytecode = True  # the reference directory is read-only
        for name in ("pyopencl", "pyopencl.array", "pygame", "siphash24"):
            if name not in sys.modules:
                sys.modules[name] = types.ModuleType(name)
        sys.modules["pyopencl"].array = sys.modules["pyopencl.array"]
        if d not in sys.path:
            sys.path.insert(0, d)
        self.particles = importlib.import_module("particles")
        self.decay_chains = importlib.import_module("decay_chains")
        self.nuclear_forces = importlib.import_module("nuclear_forces")
        self.dir = d

    def nuclear_sim(self):
        """The reference's application module (imports rendering -> the pygame stub)."""
        return importlib.import_module("nuclear_sim")

    def forces(self, strong=150.0, coulomb=30.0, pauli=35.0):
        nf = object.__new__(self.nuclear_forces.NuclearForces)
        nf.strong_strength = strong
        nf.coulomb_strength = coulomb
        nf.pauli_strength = pauli
        nf.gravity_strength = 0.01
        nf.weak_strength = 1.0
        return nf

    def make_particles(self, x, y, vx, vy, is_proton):
        P, T = self.particles.Particle, self.particles.ParticleType
        return [P(float(a), float(b), T.PROTON if t else T.NEUTRON, float(c), float(d))
                for a, b, c, d, t in zip(x, y, vx, vy, is_proton)]


class DrawFeeder:
    """Stand-in for the ``random`` module inside a reference module: serves the
    supplied uniforms in order (``random()``), and derives ``uniform``/``randint`` from them
    exactly like CPython does (``a + (b-a)*random()``)."""

    def __init__(self, draws):
        self.draws = list(draws)
        self.used = 0

    def random(self):
        u = self.draws[self.used]
        self.used += 1
        return u

    def uniform(self, a, b):
        return a + (b - a) * self.random()

    def randint(self, a, b):
        return a + int(self.random() * (b - a + 1))
