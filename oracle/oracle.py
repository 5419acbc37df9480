"""ctypes/numpy front end of the CPU oracle (oracle/pyqmd_oracle.c).

TEST INFRASTRUCTURE ONLY -- the checker for the CUDA path, never the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module.  See the header of ``pyqmd_oracle.c`` for the parity status
(pinned against outputs of the unmodified reference, fixtures under ``tests/golden/``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


class BranchStats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in
                ("evaluated", "skipped", "hard", "core", "attr", "tail", "pp", "pauli",
                 "clamped", "contained")]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}

    def flops(self):
        """Algorithmic FLOPs by the convention of SURVEY.md section 8(d): 15 per evaluated pair,
        + 4 / 7 / 8 for the core / attractive / tail strong branch, + 5 hard core,
        + 3 proton-proton, + 6 Pauli."""
        return (15 * self.evaluated + 4 * self.core + 7 * self.attr + 8 * self.tail +
                5 * self.hard + 3 * self.pp + 6 * self.pauli)


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile if missing/stale."""
    src = os.path.join(_HERE, "pyqmd_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_force_step.argtypes = [C.c_int64, _f64p, _f64p, _f64p, _f64p, _u8p, C.c_double,
                                     C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_double, C.c_void_p, C.c_int]
        L.orc_force_step.restype = None
        L.orc_ensemble_force_steps.argtypes = [C.c_int64, _i64p, _i32p, _f64p, _f64p, _f64p, _f64p,
                                               _u8p, C.c_double, C.c_double, C.c_double,
                                               C.c_double, C.c_int, C.c_int]
        L.orc_ensemble_force_steps.restype = C.c_int64
        L.orc_cloud_forces.argtypes = [C.c_int64, _f64p, _f64p, _u8p, C.c_double, C.c_double,
                                       C.c_double, C.c_double, C.c_double, C.c_int64, C.c_int64,
                                       _f64p, _f64p, C.c_int]
        L.orc_cloud_forces.restype = None
        L.orc_resolve_overlaps.argtypes = [C.c_int64, _f64p, _f64p, C.c_void_p, C.c_int64,
                                           C.POINTER(C.c_int64)]
        L.orc_resolve_overlaps.restype = C.c_int64
        L.orc_decay_probability.argtypes = [C.c_double, C.c_double]
        L.orc_decay_probability.restype = C.c_double
        L.orc_decay_decisions.argtypes = [C.c_int64, _f64p, C.c_double, _f64p, _u8p, C.c_void_p,
                                          C.c_int]
        L.orc_decay_decisions.restype = C.c_int64
        L.orc_py312_mean.argtypes = [_f64p, C.c_int64]
        L.orc_py312_mean.restype = C.c_double
        L.orc_u53.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_u53.restype = C.c_double
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                        C.POINTER(C.c_uint32)]
        L.orc_philox4x32_10.restype = None
        L.orc_philox_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        L.orc_philox_uniform.restype = C.c_double
        L.orc_philox_uniforms.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_uint32,
                                          C.c_uint32, _f64p]
        L.orc_philox_uniforms.restype = None
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def force_step(x, y, vx, vy, is_proton, dt, S=150.0, Cc=30.0, P=35.0, *, integrate=True,
               amb_tol=0.0, want_forces=False, want_stats=False):
    """One reference step (nuclear_forces.py:236-323) on float64 arrays, in place.

    Returns a dict with optional 'fx','fy' (pre-integration forces), 'amb' (threshold
    ambiguity flags at relative ``amb_tol``) and 'stats' (BranchStats)."""
    n = len(x)
    for a in (x, y, vx, vy):
        assert a.dtype == np.float64 and a.flags.c_contiguous and len(a) == n
    t = np.ascontiguousarray(is_proton, dtype=np.uint8)
    fx = np.empty(n) if want_forces else None
    fy = np.empty(n) if want_forces else None
    amb = np.zeros(n, np.uint8) if amb_tol > 0 else None
    st = BranchStats() if want_stats else None
    lib().orc_force_step(n, x, y, vx, vy, t, S, Cc, P, dt, _p(fx), _p(fy), _p(amb), amb_tol,
                         C.cast(C.pointer(st), C.c_void_p) if st is not None else None,
                         1 if integrate else 0)
    out = {}
    if want_forces:
        out["fx"], out["fy"] = fx, fy
    if amb is not None:
        out["amb"] = amb.astype(bool)
    if st is not None:
        out["stats"] = st
    return out


def ensemble_force_steps(offsets, count, x, y, vx, vy, is_proton, dt, n_steps, S=150.0, Cc=30.0,
                         P=35.0, n_threads=0):
    """n_steps reference steps of every nucleus of a CSR-packed ensemble (OpenMP)."""
    offsets = np.ascontiguousarray(offsets, np.int64)
    count = np.ascontiguousarray(count, np.int32)
    t = np.ascontiguousarray(is_proton, np.uint8)
    return int(lib().orc_ensemble_force_steps(len(count), offsets, count, x, y, vx, vy, t, S, Cc,
                                              P, dt, n_steps, n_threads))


def cloud_forces(x, y, is_proton, i0, i1, S=150.0, Cc=30.0, P=35.0, center=None, n_threads=0):
    """Forces (incl. containment) on nucleons [i0, i1) of one big cloud (OpenMP over i)."""
    n = len(x)
    t = np.ascontiguousarray(is_proton, np.uint8)
    if center is None:
        center = (py312_mean(x), py312_mean(y))
    fx = np.empty(i1 - i0)
    fy = np.empty(i1 - i0)
    lib().orc_cloud_forces(n, x, y, t, S, Cc, P, center[0], center[1], i0, i1, fx, fy, n_threads)
    return fx, fy


def resolve_overlaps(x, y, uniforms=()):
    """nuclear_sim.py:355-379 on float64 arrays, in place.  Returns (draws consumed, pushes)."""
    u = np.ascontiguousarray(uniforms, np.float64)
    pushes = C.c_int64(0)
    used = lib().orc_resolve_overlaps(len(x), x, y, _p(u) if len(u) else None, len(u),
                                      C.byref(pushes))
    if used < 0:
        raise ValueError("resolve_overlaps needed more uniforms than supplied")
    return int(used), int(pushes.value)


def decay_probability(T, dt):
    """particles.py:126-144; -1.0 means stable (no draw is consumed)."""
    return float(lib().orc_decay_probability(float(T), float(dt)))


def decay_decisions(T, dt, u, n_threads=0):
    T = np.ascontiguousarray(T, np.float64)
    u = np.ascontiguousarray(u, np.float64)
    out = np.empty(len(T), np.uint8)
    consumed = np.empty(len(T), np.uint8)
    lib().orc_decay_decisions(len(T), T, float(dt), u, out, _p(consumed), n_threads)
    return out.astype(bool), consumed.astype(bool)


def py312_mean(v):
    v = np.ascontiguousarray(v, np.float64)
    return float(lib().orc_py312_mean(v, len(v)))


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*[int(v) & 0xFFFFFFFF for v in ctr])
    k = (C.c_uint32 * 2)(*[int(v) & 0xFFFFFFFF for v in key])
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(v) for v in o]


def philox_uniform(seed, nucleus_id, step, slot):
    return float(lib().orc_philox_uniform(seed, nucleus_id, step, slot))


def philox_uniforms(seed, id0, n, step, slot):
    out = np.empty(n)
    lib().orc_philox_uniforms(seed, id0, n, step, slot, out)
    return out


def u53(w0, w1):
    return float(lib().orc_u53(w0, w1))


def max_threads():
    return int(lib().orc_max_threads())
