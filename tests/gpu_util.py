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


# ---- threshold-ambiguous nucleons: the device must have taken one of the admissible branches ------------
THRESHOLDS = (0.1, 2.8, 4.25, 8.0, 9.0)      # d at which the law jumps (d2 = 0.01, :257; :264,273,276,289)


def pair_net(d, ti, tj, S=150.0, C=30.0, P=35.0):
    """net(d) of one pair, nuclear_forces.py:257-294 (test-side numpy restatement; the C oracle is the
    reference for sums, this is only used to build the alternatives at a discontinuity)."""
    if d * d < 0.01:
        return None                                          # skipped pair (:257)
    net = 0.0
    if d < 4.25:
        net -= 60.0 * ((4.25 - d) / 4.25) ** 1.5
    if d < 2.8:
        net -= 0.7 * S / (d * d + 0.15)
    elif d < 9.0:
        net += 1.25 * S * np.exp(-d / 7.0) / (d + 0.15)
    else:
        net += 0.15 * S * np.exp(-1.8 * d / 7.0) / (d + 0.15)
    if ti and tj:
        net -= C / (d * d + 0.15)
    if ti == tj and d < 8.0:
        net -= P * np.exp(-2.0 * d / 8.0)
    return max(-12.0, min(12.0, net))


def admissible_force_check(pos32, isp, f_dev, fx, fy, amb, S=150.0, C=30.0, P=35.0, tol=AMB_TOL):
    """For every nucleon the oracle flagged as ambiguous: the device force must equal the oracle force
    with each near-threshold pair taken on EITHER side of its discontinuity (all 2^k combinations of the
    k ambiguous pairs, k small).  Returns (checked, worst relative mismatch of the best combination)."""
    import itertools
    p = pos32.astype(np.float64)
    worst, checked = 0.0, 0
    for i in np.nonzero(amb)[0]:
        dx, dy = p[:, 0] - p[i, 0], p[:, 1] - p[i, 1]
        d = np.hypot(dx, dy)
        alts = []                                            # (delta_fx, delta_fy) of flipping pair (i, j)
        thr_a = np.array(THRESHOLDS)
        near = np.abs(d[:, None] - thr_a[None, :]) <= 4 * tol * thr_a[None, :]
        near[i] = False
        for j, t in zip(*np.nonzero(near)):
            thr = THRESHOLDS[t]
            cur = pair_net(d[j], isp[i], isp[j], S, C, P)
            other = pair_net(thr * (1 + 8 * tol) if d[j] < thr else thr * (1 - 8 * tol), isp[i], isp[j], S, C, P)
            c = 0.0 if cur is None else cur
            o = 0.0 if other is None else other
            alts.append(((o - c) * dx[j] / d[j], (o - c) * dy[j] / d[j]))
        if not alts or len(alts) > 6:
            continue
        best = np.inf
        for mask in itertools.product((0, 1), repeat=len(alts)):
            ex = fx[i] + sum(m * a[0] for m, a in zip(mask, alts))
            ey = fy[i] + sum(m * a[1] for m, a in zip(mask, alts))
            best = min(best, np.hypot(f_dev[i, 0] - ex, f_dev[i, 1] - ey) / max(np.hypot(ex, ey), 1.0))
        worst = max(worst, best)
        checked += 1
    return checked, worst
