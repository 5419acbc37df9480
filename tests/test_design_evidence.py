"""Design evidence that needs no GPU: numbers DESIGN.md quotes for choices that were NOT (yet) taken.

VERDICT r01 (item 2) proposed a warp-uniform far path for settled nuclei in the ensemble kernel: order
the nucleons of a nucleus along a space-filling curve and let a warp vote, per ring step, on "all my
pairs are >= 9 apart" to take the 2-MUFU tail-only law (as cloud_sym_kernel does per 256-nucleon
tile).  This test measures on settled Pb-208 nuclei (reference layout, 12 app frames of the oracle:
4 sub-steps + resolve_overlaps) how often such a vote would pass with the ring kernel's granularity
(4 nucleons per lane, rings of 26 lanes, 2 warps), for Morton-ordered nucleons and for the reference's
list order (shell by shell, which is what the kernels sweep today).

Outcome on the B200 (round 2, profiles/raw_r02/far_vote_ab.log): with the vote in both ensemble kernels
(per packed call, `__all_sync` on d2 >= 81, far law = 2 MUFU + 9 FMA-pipe ops) every workload lost 15-23 %
-- Pb-208 free-running 1.147e12 -> 0.973e12 pairs/s, U-238 1.206e12 -> 0.930e12 -- because a branch
per call stops the compiler from interleaving the eight independent evaluations of a ring step, which is
what hides the MUFU latency; and settled, Morton-ordered nuclei (39 % of the calls all-far, below) were
no faster than list-ordered ones (0.971e12 vs 0.974e12).  Not shipped; the numbers stay here.
"""
import numpy as np

from oracle import oracle as orc
from pyqmd_b200.state import layout_templates


def morton_order(x, y):
    def spread(v):
        v = v.astype(np.uint64)
        out = np.zeros_like(v)
        for b in range(16):
            out |= ((v >> np.uint64(b)) & np.uint64(1)) << np.uint64(2 * b)
        return out
    qx = ((x - x.min()) / (np.ptp(x) + 1e-9) * 65535).astype(np.uint64)
    qy = ((y - y.min()) / (np.ptp(y) + 1e-9) * 65535).astype(np.uint64)
    return np.argsort(spread(qx) | (spread(qy) << np.uint64(1)), kind="stable")


def settled_pb208(template):
    tm = layout_templates()
    xy = tm["z82_n126_xy"][template].astype(np.float64)
    isp = tm["z82_n126_isp"][template]
    x, y = xy[:, 0].copy(), xy[:, 1].copy()
    vx, vy = np.zeros_like(x), np.zeros_like(x)
    u = np.random.default_rng(template).random(4096)
    for _ in range(12):
        for _ in range(4):
            orc.force_step(x, y, vx, vy, isp, 1 / 240)
        orc.resolve_overlaps(x, y, u)
    return x, y


def vote_pass_rates(x, y, order):
    """(whole ring steps, packed calls = one i of every lane against one j pair) that are all-far."""
    n = len(x)
    d = np.hypot(x[:, None] - x[None, :], y[:, None] - y[None, :])
    sub = [order[4 * k:4 * k + 4] for k in range(n // 4)]          # 52 subgroups
    P = 26
    steps = step_ok = calls = call_ok = 0
    for gi in range(2):
        for gj in range(2):
            for m in range(P):                          # ring step: lane l meets subgroup (l + m) mod P
                if gi == gj and m == 0:
                    continue
                blk = np.array([[d[np.ix_(sub[gi * P + l], sub[gj * P + (l + m) % P])] for l in range(P)]])[0]
                far = blk >= 9.0                        # [lane, i, j]
                steps += 1
                step_ok += bool(far.all())
                for i in range(4):
                    for jp in range(2):
                        calls += 1
                        call_ok += bool(far[:, i, 2 * jp:2 * jp + 2].all())
    return step_ok / steps, call_ok / calls


def test_how_often_a_warp_uniform_far_vote_would_pass_inside_a_nucleus():
    far_pairs, morton, listed = [], [], []
    for t in range(6):
        x, y = settled_pb208(t)
        n = len(x)
        d = np.hypot(x[:, None] - x[None, :], y[:, None] - y[None, :])
        far_pairs.append(float((d[np.triu_indices(n, 1)] >= 9.0).mean()))
        morton.append(vote_pass_rates(x, y, morton_order(x, y)))
        listed.append(vote_pass_rates(x, y, np.arange(n)))
    frac_pairs = float(np.mean(far_pairs))
    m_step, m_call = np.mean(morton, 0)
    l_step, l_call = np.mean(listed, 0)
    print(f"settled Pb-208: {frac_pairs:.1%} of the pairs are >= 9 apart; all-far ring steps / packed calls: "
          f"Morton order {m_step:.1%} / {m_call:.1%}, reference list order {l_step:.1%} / {l_call:.1%}")
    assert frac_pairs > 0.85                    # the tail branch dominates a settled nucleus
    assert l_call < 0.10                        # ... but in list order a warp-uniform vote hardly ever passes
    assert 0.05 < m_step < 0.5 and m_call > m_step      # Morton order: a minority of the steps
