"""ctypes binding of libpyqmd_b200.so (C ABI declared in include/pyqmd_b200.h).

The product path has no CPU fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.  Build it with ``python -m pyqmd_b200.build`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("PYQMD_B200_LIB") or os.path.join(HERE, "libpyqmd_b200.so")   # override: tuning builds

TABLE_ZDIM = 128
TABLE_NDIM = 192
COUNT_COLS = 16
HL_INF, HL_TABLE, HL_BAND = 0, 1, 2

# every symbol include/pyqmd_b200.h declares
EXPORTS = (
    "pyqmd_abi_version", "pyqmd_last_error", "pyqmd_device_props", "pyqmd_struct_sizes",
    "pyqmd_fp32_peak",
    "pyqmd_update_forces_and_positions", "pyqmd_update_particles_f64", "pyqmd_cloud_step_host",
    "pyqmd_cloud_workspace_bytes", "pyqmd_cloud_step", "pyqmd_cloud_sort_keys",
    "pyqmd_cloud_force_scale_log2", "pyqmd_cloud_pair_forces", "pyqmd_cloud_pair_forces_ex", "pyqmd_cloud_integrate",
    "pyqmd_cloud_exchange_integrate",
    "pyqmd_ensemble_step", "pyqmd_ensemble_step_host", "pyqmd_resolve_overlaps", "pyqmd_ensemble_census",
    "pyqmd_ensemble_init_layout",
    "pyqmd_population_step", "pyqmd_free_particles_frame",
)

# numpy mirror of pyqmd_nuclide_entry (88 bytes)
NUCLIDE_DTYPE = np.dtype([
    ("half_life", "<f8"), ("p_decay", "<f8"), ("band_a", "<f8"), ("band_b", "<f8"),
    ("band_unit", "<f8"), ("opt_cum", "<f8", (2,)), ("opt_zn", "<i4", (2,)),
    ("opt_mode", "<i4", (2,)), ("n_opt", "<i4"), ("kind", "<i4"), ("p_thr", "<u8"),
], align=True)
THR_PER_NUCLEUS = 0xFFFFFFFFFFFFFFFF

# numpy mirror of pyqmd_decay_event (56 bytes)
EVENT_DTYPE = np.dtype([
    ("nucleus", "<i8"), ("step", "<i4"), ("mode", "<i4"), ("zn_new", "<i4"), ("ptype", "<i4"),
    ("x", "<f8"), ("y", "<f8"), ("vx", "<f8"), ("vy", "<f8"),
], align=True)
# numpy mirror of pyqmd_free_particle (64 bytes)
FREE_DTYPE = np.dtype([
    ("x", "<f8"), ("y", "<f8"), ("vx", "<f8"), ("vy", "<f8"), ("age", "<f8"), ("lifetime", "<f8"),
    ("nucleus", "<i8"), ("type", "<i4"), ("pad", "<i4"),
], align=True)
assert NUCLIDE_DTYPE.itemsize == 88 and EVENT_DTYPE.itemsize == 56 and FREE_DTYPE.itemsize == 64


class EnsembleDesc(C.Structure):
    """pyqmd_ensemble"""
    _fields_ = [
        ("pos", C.c_void_p), ("vel", C.c_void_p), ("is_proton", C.c_void_p),
        ("force", C.c_void_p), ("offset", C.c_void_p), ("count", C.c_void_p),
        ("zn", C.c_void_p), ("half_life", C.c_void_p), ("p_decay", C.c_void_p),
        ("origin", C.c_void_p), ("centre", C.c_void_p),
        ("n_nuclei", C.c_int64), ("id_base", C.c_int64),
        ("list", C.c_void_p), ("n_list", C.c_int64),
        ("cap", C.c_int32), ("decay_enabled", C.c_int32),
        ("strong", C.c_float), ("coulomb", C.c_float), ("pauli", C.c_float),
        ("dt_phys", C.c_float), ("dt_decay", C.c_double),
        ("table", C.c_void_p), ("uniforms", C.c_void_p), ("uniforms_n", C.c_int64),
        ("seed", C.c_uint64), ("step0", C.c_uint32), ("reserved", C.c_uint32),
        ("events", C.c_void_p), ("event_capacity", C.c_int64), ("event_count", C.c_void_p),
        ("mode_counts", C.c_void_p),
    ]


MAX_CHUNK_LAUNCHES = 12


class HostChunk(C.Structure):
    """pyqmd_host_chunk"""
    _fields_ = [
        ("nuc0", C.c_int64), ("nuc1", C.c_int64), ("slot0", C.c_int64), ("slot1", C.c_int64),
        ("n_launch", C.c_int32), ("cap", C.c_int32 * MAX_CHUNK_LAUNCHES),
        ("list", C.c_void_p * MAX_CHUNK_LAUNCHES), ("n_list", C.c_int64 * MAX_CHUNK_LAUNCHES),
    ]


class PopulationDesc(C.Structure):
    """pyqmd_population"""
    _fields_ = [
        ("zn", C.c_void_p), ("half_life", C.c_void_p), ("p_decay", C.c_void_p),
        ("n", C.c_int64), ("id_base", C.c_int64), ("table", C.c_void_p),
        ("dt_decay", C.c_double), ("uniforms", C.c_void_p), ("uniforms_n", C.c_int64),
        ("seed", C.c_uint64), ("step0", C.c_uint32), ("n_watch", C.c_int32),
        ("watch_zn", C.c_int32 * 8), ("step_counts", C.c_void_p), ("decided", C.c_void_p),
        ("flags", C.c_int32), ("reserved", C.c_int32),
    ]


POP_PER_NUCLEUS_STATE = 1
CLOUD_SKIP_EXACT_ZEROS = 1


class FreeFrame(C.Structure):
    """pyqmd_free_frame"""
    _fields_ = [
        ("num_steps", C.c_int32), ("step0", C.c_uint32), ("fast_forward", C.c_int32), ("reserved", C.c_int32),
        ("speed_scale", C.c_double), ("aging_scale", C.c_double), ("age_dt", C.c_double),
        ("nucleon_dt", C.c_double), ("lifetime_fast", C.c_double), ("lifetime_floor", C.c_double),
    ]


_lib = None


def lib():
    """Load libpyqmd_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            f"{SO_PATH} not found: the CUDA extension is required (no CPU fallback). "
            "Build it with `python -m pyqmd_b200.build`.")
    L = C.CDLL(SO_PATH)
    vp, i64, i32, f32, f64 = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_double
    L.pyqmd_abi_version.restype = C.c_int
    L.pyqmd_last_error.restype = C.c_char_p
    L.pyqmd_device_props.argtypes = [C.c_int, C.POINTER(i64)]
    L.pyqmd_struct_sizes.argtypes = [C.POINTER(i64)]
    L.pyqmd_fp32_peak.argtypes = [C.c_int, C.POINTER(f64), C.POINTER(f64), vp]
    L.pyqmd_update_forces_and_positions.argtypes = [vp, vp, i32, f32, f32, f32, f32, f32, f32]
    L.pyqmd_update_particles_f64.argtypes = [vp, vp, vp, vp, vp, i64, f64, f64, f64, f64, i32]
    L.pyqmd_cloud_step_host.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, f32, i32]
    L.pyqmd_cloud_workspace_bytes.argtypes = [i64]
    L.pyqmd_cloud_workspace_bytes.restype = i64
    L.pyqmd_cloud_step.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, f32, f32, f32, f32, vp, vp]
    L.pyqmd_cloud_sort_keys.argtypes = [vp, vp, i64, f32, f32, f32, vp, vp]
    L.pyqmd_cloud_force_scale_log2.argtypes = [i64]
    L.pyqmd_cloud_pair_forces.argtypes = [vp, vp, i64, i32, i32, f32, f32, f32, vp, vp, vp]
    L.pyqmd_cloud_pair_forces_ex.argtypes = [vp, vp, i64, i32, i32, f32, f32, f32, vp, vp, C.c_uint32, vp]
    L.pyqmd_cloud_integrate.argtypes = [vp, vp, vp, vp, i64, i64, i64, f32, vp, vp, vp]
    L.pyqmd_cloud_exchange_integrate.argtypes = [vp, vp, vp, i64, i64, i64, f32, vp, vp, i32, vp, vp]
    L.pyqmd_ensemble_step.argtypes = [C.POINTER(EnsembleDesc), i32, vp]
    L.pyqmd_ensemble_step_host.argtypes = [C.POINTER(EnsembleDesc), vp, vp, vp, vp, vp,
                                           C.POINTER(HostChunk), i32, i32, vp]
    L.pyqmd_resolve_overlaps.argtypes = [C.POINTER(EnsembleDesc), vp, i32, vp, vp]
    L.pyqmd_ensemble_census.argtypes = [C.POINTER(EnsembleDesc), vp, vp]
    L.pyqmd_ensemble_init_layout.argtypes = [C.POINTER(EnsembleDesc), C.POINTER(C.c_double), vp,
                                             C.c_uint64, vp]
    L.pyqmd_population_step.argtypes = [C.POINTER(PopulationDesc), i32, vp]
    L.pyqmd_free_particles_frame.argtypes = [vp, vp, vp, vp, i64, vp, vp, i64, C.POINTER(FreeFrame), vp, i32, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("pyqmd_last_error", "pyqmd_cloud_workspace_bytes"):
            fn.restype = C.c_int
    _lib = L
    return L


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().pyqmd_last_error().decode(errors="replace")
        raise RuntimeError(f"libpyqmd_b200 {what} failed (code {rc}): {msg}")


def require_cuda():
    """The product path runs on a CUDA device only; fail loudly otherwise."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("pyqmd_b200 needs a CUDA device (B200, sm_100a); there is no CPU "
                           "fallback on the product path")
    lib()


def ptr(t):
    """Device (or host) pointer of a torch tensor / numpy array, None -> NULL."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
