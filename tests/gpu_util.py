"""Shared helpers of the GPU parity tests: tolerance norms of SURVEY.md section 8(d)."""
import numpy as np

from oracle import oracle as orc

POS_TOL = 1e-5      # max_i |dx_i| / max_i |x_i - x_cm|   per step, from identical FP32 state
FORCE_TOL = 1e-5    # ||dF||_2 / ||F||_2                  per step
AMB_TOL = 1e-6      # pairs this close (relative) to a branch threshold may flip in FP32


def oracle_step(pos32, vel32, isp, dt, S=150.0, C=30.0, P=35.0, integrate=True):
    """Reference step (float64) from an FP32-representable state.  Returns new x,y,vx,vy,
    forces and the per-nucleon ambiguity mask."""
    x = pos32[:, 0].astype(np.float64).copy()
    y = pos32[:, 1].astype(np.float64).copy()
    vx = vel32[:, 0].astype(np.float64).copy()
    vy = vel32[:, 1].astype(np.float64).copy()
    r = orc.force_step(x, y, vx, vy, isp, dt, S, C, P, integrate=integrate, amb_tol=AMB_TOL,
                       want_forces=True)
    return x, y, vx, vy, r["fx"], r["fy"], r["amb"]


def extent_of(pos32):
    p = pos32.astype(np.float64)
    c = p.mean(0)
    return max(float(np.sqrt(((p - c) ** 2).sum(1)).max()), 1e-3)


def pos_error(pos32_before, new_pos_dev, ox, oy, amb):
    """max_i ||x_dev - x_oracle|| / extent over unambiguous nucleons."""
    ok = ~amb
    if not ok.any():
        return 0.0
    d = np.hypot(new_pos_dev[:, 0].astype(np.float64) - ox, new_pos_dev[:, 1].astype(np.float64) - oy)
    return float(d[ok].max() / extent_of(pos32_before))


def force_error(f_dev, fx, fy, amb):
    ok = ~amb
    if not ok.any():
        return 0.0
    dfx = f_dev[ok, 0].astype(np.float64) - fx[ok]
    dfy = f_dev[ok, 1].astype(np.float64) - fy[ok]
    num = np.sqrt((dfx ** 2 + dfy ** 2).sum())
    den = np.sqrt((fx[ok] ** 2 + fy[ok] ** 2).sum())
    return float(num / max(den, 1e-30))


def single_nucleus_ensemble(pos32, vel32, isp, **kw):
    from pyqmd_b200.state import NucleusEnsemble
    n = len(isp)
    z = int(np.sum(isp))
    zn = np.array([(z << 16) | (n - z)], np.int32)
    kw.setdefault("decay", False)
    return NucleusEnsemble(zn, np.array([0], np.int64), np.array([n], np.int32), pos32, vel32,
                           isp, keep_force=True, **kw)
