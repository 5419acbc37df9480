"""Drop-in for the reference's ``nuclear_forces.NuclearForces`` (OtsoBear/PyQMD
nuclear_forces.py:10-323) backed by libpyqmd_b200.so.

Same constructor (no arguments), same attributes (``strong_strength`` ... ``weak_strength``,
nuclear_forces.py:13-17) and the same two methods the app calls (nuclear_sim.py:171,173):

    update_particles_gpu(particles, dt)   # list[Particle], mutated in place, returns None
    update_particles_cpu(particles, dt)

Both run on the GPU here (there is no CPU path in this package); ``_cpu`` keeps the
reference's float64 attribute types, ``_gpu`` mirrors the OpenCL path's float32 buffers
(h_particles / h_types / center, nuclear_forces.py:190-219).  Differences from the reference,
all deliberate: the step is Jacobi like the reference's CPU path (its OpenCL kernel updates in
place and races, :168-171); construction needs no OpenCL; an empty list returns immediately
(:186-188, :238-239); a failing kernel raises instead of being logged and ignored (:222-224).
"""
from __future__ import annotations

import ctypes as C
import logging

import numpy as np

from . import _lib
from .types import ParticleType

logger = logging.getLogger("NuclearSim")      # same logger name as the reference (:8)

_PROTON = ParticleType.PROTON.value           # 0, particles.py:6


def _is_proton(p):
    """The caller's particles may be the reference's own ``particles.Particle`` objects (nuclear_sim.py
    builds the nucleus with ITS ParticleType enum when only nuclear_forces is swapped): compare the
    enum VALUE, an ``==`` between two different Enum classes is silently False for every nucleon."""
    t = p.type
    return getattr(t, "value", t) == _PROTON


class NuclearForces:
    def __init__(self):
        self.setup_device()
        self.strong_strength = 150.0           # nuclear_forces.py:13
        self.coulomb_strength = 30.0           # :14
        self.pauli_strength = 35.0             # :15
        self.gravity_strength = 0.01           # :16 (unused by the reference as well)
        self.weak_strength = 1.0               # :17 (unused by the reference as well)

    def setup_device(self):
        """Counterpart of setup_opencl (:19-54): raises RuntimeError when no device / library,
        which nuclear_sim.py:40-45 turns into ``gpu_available = False``."""
        lib = _lib.lib()
        props = (C.c_int64 * 8)()
        rc = lib.pyqmd_device_props(0, props)
        if rc != 0:
            raise RuntimeError("No CUDA device found: " + lib.pyqmd_last_error().decode())
        self.device_props = dict(sm_count=int(props[0]), cc=(int(props[1]), int(props[2])),
                                 sm_clock_khz=int(props[3]), l2_bytes=int(props[4]))
        logger.info("Using GPU: %d SMs, compute capability %d.%d", props[0], props[1], props[2])

    # -- reference-shaped calls ---------------------------------------------------------------------
    def update_particles_gpu(self, particles, dt):
        """nuclear_forces.py:185-234 with the OpenCL launch replaced by
        pyqmd_update_forces_and_positions; float32 buffers exactly as the reference packs them."""
        n = len(particles)
        if n == 0:
            return
        h_particles = np.zeros((n, 4), dtype=np.float32)                     # :190
        h_types = np.zeros(n, dtype=np.int32)                                # :191
        for i, p in enumerate(particles):                                    # :194-199
            h_particles[i] = (p.x, p.y, p.vx, p.vy)
            h_types[i] = 0 if _is_proton(p) else 1
        center_x = sum(p.x for p in particles) / n                           # :206-207
        center_y = sum(p.y for p in particles) / n
        if n > 1024:
            return self._update_f64(particles, dt)
        rc = _lib.lib().pyqmd_update_forces_and_positions(
            h_particles.ctypes.data, h_types.ctypes.data, n, center_x, center_y,
            self.strong_strength, self.coulomb_strength, self.pauli_strength, dt)
        _lib.check(rc, "pyqmd_update_forces_and_positions")
        for i, p in enumerate(particles):                                    # :230-234
            p.x, p.y, p.vx, p.vy = h_particles[i]

    def update_particles_cpu(self, particles, dt):
        """Same contract as nuclear_forces.py:236-323 (float64 in, float64 out); computed on
        the GPU in nucleus-relative FP32."""
        if not particles:
            return
        self._update_f64(particles, dt)

    def step(self, particles, dt, n_steps):
        """``n_steps`` consecutive sub-steps (the loop nuclear_sim.py:161-173 without decay)
        with a single host<->device round trip."""
        if particles and n_steps > 0:
            self._update_f64(particles, dt, n_steps)

    def step_arrays(self, x, y, vx, vy, is_proton, dt, n_steps=1):
        """The same update on caller-owned float64 numpy arrays (updated in place), for callers that
        keep their state in arrays rather than in ``Particle`` objects: no per-object marshalling."""
        n = len(x)
        if n == 0 or n_steps <= 0:
            return
        for a in (x, y, vx, vy):
            if a.dtype != np.float64 or not a.flags.c_contiguous or len(a) != n:
                raise ValueError("x, y, vx, vy must be C-contiguous float64 arrays of equal length")
        isp = np.ascontiguousarray(is_proton, dtype=np.uint8)
        rc = _lib.lib().pyqmd_update_particles_f64(
            x.ctypes.data, y.ctypes.data, vx.ctypes.data, vy.ctypes.data, isp.ctypes.data, n,
            self.strong_strength, self.coulomb_strength, self.pauli_strength, dt, n_steps)
        _lib.check(rc, "pyqmd_update_particles_f64")

    def step_cloud(self, pos, vel, is_proton, dt, n_steps=1, force=None):
        """One large system in caller-owned float32 host arrays ``pos[n, 2]``, ``vel[n, 2]`` (numpy or
        pinned torch CPU tensors, updated in place) and ``is_proton[n]`` (uint8): the reference's
        per-step call (nuclear_forces.py:185-234) at any N through pyqmd_cloud_step_host -- upload,
        sort once, ``n_steps`` steps of the symmetric scheme, un-sort, download.  ``force``: optional
        float32 [n, 2] array receiving the force of the last step."""
        n = int(pos.shape[0])
        if n == 0 or n_steps <= 0:
            return
        rc = _lib.lib().pyqmd_cloud_step_host(
            _lib.ptr(pos), _lib.ptr(vel), _lib.ptr(is_proton), _lib.ptr(force), n,
            self.strong_strength, self.coulomb_strength, self.pauli_strength, dt, n_steps)
        _lib.check(rc, "pyqmd_cloud_step_host")

    def _update_f64(self, particles, dt, n_steps=1):
        n = len(particles)
        x = np.fromiter((p.x for p in particles), np.float64, n)
        y = np.fromiter((p.y for p in particles), np.float64, n)
        vx = np.fromiter((p.vx for p in particles), np.float64, n)
        vy = np.fromiter((p.vy for p in particles), np.float64, n)
        isp = np.fromiter((_is_proton(p) for p in particles), np.uint8, n)
        rc = _lib.lib().pyqmd_update_particles_f64(
            x.ctypes.data, y.ctypes.data, vx.ctypes.data, vy.ctypes.data, isp.ctypes.data, n,
            self.strong_strength, self.coulomb_strength, self.pauli_strength, dt, n_steps)
        _lib.check(rc, "pyqmd_update_particles_f64")
        for p, a, b, c, d in zip(particles, x.tolist(), y.tolist(), vx.tolist(), vy.tolist()):
            p.x, p.y, p.vx, p.vy = a, b, c, d
