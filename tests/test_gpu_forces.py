"""GPU parity of the force + integrate path against the oracle (which is pinned bit-exact to the
reference, tests/test_oracle.py).  Every call goes through the C ABI of libpyqmd_b200.so.

Tolerances, per step from an identical FP32-representable state:
  max_i |dx_i| / max_i |x_i - x_cm| <= 1e-5   and   ||dF||_2 / ||F||_2 <= 1e-5,
nucleons with a pair within 1e-6 (relative) of a branch threshold are counted separately
(an FP32 evaluation may legitimately take the other branch there).
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, unhex
from gpu_util import (AMB_TOL, FORCE_TOL, POS_TOL, admissible_force_check, extent_of, force_error, oracle_step,
                      pos_error,
                      single_nucleus_ensemble)
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _particles(case_state):
    from pyqmd_b200 import Particle, ParticleType
    return [Particle(x, y, ParticleType.PROTON if t else ParticleType.NEUTRON, vx, vy)
            for x, y, vx, vy, t in zip(unhex(case_state["x"]), unhex(case_state["y"]),
                                       unhex(case_state["vx"]), unhex(case_state["vy"]),
                                       case_state["is_proton"])]


@pytest.mark.parametrize("method", ["update_particles_cpu", "update_particles_gpu"])
def test_known_answer_cases(force_kats, method, ensemble_kernel):
    """Reference golden vectors through the drop-in NuclearForces methods."""
    from pyqmd_b200 import NuclearForces
    nf = NuclearForces()
    checked = 0
    for case in force_kats["cases"]:
        nf.strong_strength = float.fromhex(case["S"])
        nf.coulomb_strength = float.fromhex(case["C"])
        nf.pauli_strength = float.fromhex(case["P"])
        ps = _particles(case["input"])
        dt = float.fromhex(case["dt"])
        if case["steps"] != 1:
            continue        # multi-step cases are covered teacher-forced below
        getattr(nf, method)(ps, dt)
        want = case["output"]
        wx, wy, wvx, wvy = (unhex(want[k]) for k in ("x", "y", "vx", "vy"))
        x0 = unhex(case["input"]["x"]); y0 = unhex(case["input"]["y"])
        scale = max(np.hypot(x0 - x0.mean(), y0 - y0.mean()).max(), 1.0)
        vscale = max(np.abs(np.concatenate([wvx, wvy])).max(), 1e-3)
        gx = np.array([p.x for p in ps]); gy = np.array([p.y for p in ps])
        gvx = np.array([p.vx for p in ps]); gvy = np.array([p.vy for p in ps])
        # boundary cases (2.79/2.81, 8.99/9.01 ...) are 0.1-0.4 % away from a threshold: exact branch
        assert np.abs(gx - wx).max() / scale <= POS_TOL, case["name"]
        assert np.abs(gy - wy).max() / scale <= POS_TOL, case["name"]
        assert np.abs(gvx - wvx).max() / vscale <= 2e-5, case["name"]
        assert np.abs(gvy - wvy).max() / vscale <= 2e-5, case["name"]
        checked += 1
    assert checked >= 45


def test_empty_list_is_a_no_op():
    from pyqmd_b200 import NuclearForces
    nf = NuclearForces()
    assert nf.update_particles_gpu([], 1 / 240) is None      # nuclear_forces.py:186-188
    assert nf.update_particles_cpu([], 1 / 240) is None      # :238-239


def test_gpu_method_keeps_float32_attribute_types():
    """After update_particles_gpu the reference leaves numpy.float32 scalars (:231-234)."""
    from pyqmd_b200 import NuclearForces, Particle, ParticleType
    ps = [Particle(400.0, 400.0, ParticleType.PROTON), Particle(403.0, 400.0, ParticleType.PROTON)]
    NuclearForces().update_particles_gpu(ps, 1 / 240)
    assert isinstance(ps[0].x, np.float32) and isinstance(ps[1].vx, np.float32)
    assert ps[0].vx > 0 > ps[1].vx                           # p-p at d=3: net attraction (KAT-A)


def test_u238_teacher_forced_1000_steps(u238_traj, ensemble_kernel):
    """Config C1.  Each step both sides start from the same FP32-representable state (the
    device trajectory); per-step position and force errors are gated, branch-flip candidates
    counted; results go to gpurun_out/c1_parity.json."""
    st0 = u238_traj["states"][0]
    isp = u238_traj["is_proton"].astype(np.uint8)
    dt = float(u238_traj["dt"])
    origin = st0[:, :2].mean(0)
    pos = (st0[:, :2] - origin).astype(np.float32)
    vel = st0[:, 2:].astype(np.float32)
    ens = single_nucleus_ensemble(pos, vel, isp, dt_phys=dt)
    worst_pos = worst_f = amb_worst = 0.0
    amb_total = amb_checked = 0
    for s in range(1000):
        p0 = ens.pos.cpu().numpy().copy()
        v0 = ens.vel.cpu().numpy().copy()
        ox, oy, ovx, ovy, fx, fy, amb = oracle_step(p0, v0, isp, dt)
        ens.step(1)
        p1 = ens.pos.cpu().numpy()
        f1 = ens.force.cpu().numpy()
        e_pos = pos_error(p0, p1, ox, oy, amb)
        e_f = force_error(f1, fx, fy, amb)
        worst_pos, worst_f = max(worst_pos, e_pos), max(worst_f, e_f)
        amb_total += int(amb.sum())
        assert e_pos <= POS_TOL, (s, e_pos)
        assert e_f <= FORCE_TOL, (s, e_f)
        if amb.any():
            # excluded from the norms above, but not unchecked: the device took one of the two branches
            c, w = admissible_force_check(p0, isp, f1, fx, fy, amb)
            amb_checked += c
            amb_worst = max(amb_worst, w)
            assert w <= 1e-4, (s, w)
    assert amb_checked > 0.5 * amb_total
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"c1_parity_{ensemble_kernel}.json"), "w") as f:
        json.dump(dict(config="C1 U-238 teacher-forced", kernel=ensemble_kernel, steps=1000, worst_pos_err=worst_pos,
                       worst_force_err_l2=worst_f, ambiguous_nucleon_steps=amb_total,
                       ambiguous_checked_against_branch_alternatives=amb_checked,
                       ambiguous_worst_mismatch=amb_worst, tolerance=POS_TOL), f)


def test_u238_free_running_drift_report(u238_traj):
    """1000 free-running steps on the device vs the reference trajectory: drift is REPORTED
    (the dynamics are chaotic once the nucleus has collapsed), only finiteness
    and the first steps are gated."""
    steps = list(u238_traj["steps"])
    st = u238_traj["states"]
    isp = u238_traj["is_proton"].astype(np.uint8)
    origin = st[0][:, :2].mean(0)
    pos = (st[0][:, :2] - origin).astype(np.float32)
    ens = single_nucleus_ensemble(pos, st[0][:, 2:].astype(np.float32), isp,
                                  dt_phys=float(u238_traj["dt"]))
    report = {}
    done = 0
    for s in steps[1:]:
        ens.step(int(s) - done)          # fused multi-step launches
        done = int(s)
        p = ens.pos.cpu().numpy().astype(np.float64) + origin
        ref = st[steps.index(s)][:, :2]
        drift = float(np.hypot(*(p - ref).T).max() / extent_of((ref - origin).astype(np.float32)))
        report[int(s)] = drift
        assert np.isfinite(p).all()
    assert report[1] <= POS_TOL and report[2] <= 1e-4
    with open(os.path.join(ROOT, "gpurun_out", "c1_drift.json"), "w") as f:
        json.dump(dict(config="C1 U-238 free-running drift (max |dx| / extent)", drift=report), f)


def test_multi_step_kat_cases_teacher_forced(force_kats, ensemble_kernel):
    from pyqmd_b200.state import NucleusEnsemble
    for case in force_kats["cases"]:
        if case["steps"] == 1 or len(case["input"]["x"]) < 2:
            continue
        inp = case["input"]
        pos = np.stack([unhex(inp["x"]), unhex(inp["y"])], 1).astype(np.float32)
        vel = np.stack([unhex(inp["vx"]), unhex(inp["vy"])], 1).astype(np.float32)
        isp = np.array(inp["is_proton"], np.uint8)
        dt = float.fromhex(case["dt"])
        ens = single_nucleus_ensemble(pos, vel, isp, dt_phys=dt)
        for s in range(case["steps"]):
            p0, v0 = ens.pos.cpu().numpy().copy(), ens.vel.cpu().numpy().copy()
            ox, oy, _, _, fx, fy, amb = oracle_step(p0, v0, isp, dt)
            ens.step(1)
            assert pos_error(p0, ens.pos.cpu().numpy(), ox, oy, amb) <= POS_TOL, case["name"]
            assert force_error(ens.force.cpu().numpy(), fx, fy, amb) <= FORCE_TOL, case["name"]


def test_fused_steps_equal_single_steps(ensemble_kernel):
    """n sub-steps in one launch (state kept in shared memory) == n launches of one sub-step,
    bit for bit."""
    from pyqmd_b200.state import NucleusEnsemble, README_ISOTOPES
    a = NucleusEnsemble.from_templates(README_ISOTOPES, 90, decay=False)
    b = NucleusEnsemble.from_templates(README_ISOTOPES, 90, decay=False)
    a.step(7)
    for _ in range(7):
        b.step(1)
    assert torch.equal(a.pos, b.pos) and torch.equal(a.vel, b.vel)


def test_mixed_ensemble_against_oracle(ensemble_kernel):
    """All nine preset isotopes (A = 1 ... 238), several nuclei per block for the small ones,
    three teacher-forced steps against the oracle, per-nucleus norms."""
    from pyqmd_b200.state import NucleusEnsemble, README_ISOTOPES
    ens = NucleusEnsemble.from_templates(README_ISOTOPES, 9 * 40, decay=False, keep_force=True)
    off = ens.offsets.cpu().numpy(); cnt = ens.count.cpu().numpy()
    isp = ens.is_proton.cpu().numpy()
    dt = ens.dt_phys
    worst = 0.0
    n_amb = 0
    for s in range(3):
        p0, v0 = ens.pos.cpu().numpy().copy(), ens.vel.cpu().numpy().copy()
        ens.step(1)
        p1, f1 = ens.pos.cpu().numpy(), ens.force.cpu().numpy()
        for k in range(ens.n_nuclei):
            sl = slice(off[k], off[k] + cnt[k])
            ox, oy, _, _, fx, fy, amb = oracle_step(p0[sl], v0[sl], isp[sl], dt)
            e = pos_error(p0[sl], p1[sl], ox, oy, amb)
            ef = force_error(f1[sl], fx, fy, amb)
            worst = max(worst, e)
            n_amb += int(amb.sum())
            assert e <= POS_TOL, (s, k, cnt[k], e)
            assert ef <= FORCE_TOL or np.hypot(fx, fy).max() < 1e-6, (s, k, cnt[k], ef)
    print("mixed ensemble worst pos err", worst, "ambiguous", n_amb)


def test_non_default_strengths_and_dt(ensemble_kernel):
    from pyqmd_b200.state import NucleusEnsemble
    rng = np.random.default_rng(5)
    n = 150
    pos = rng.uniform(-12, 12, (n, 2)).astype(np.float32)
    vel = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    isp = (rng.random(n) < 0.4).astype(np.uint8)
    # the last two: S = 0 and a (non-physical) negative S -- the folded log2|coef| must keep its sign
    for S, C, P, dt in ((20.0, 3.0, 4.0, 1 / 60), (900.0, 300.0, 4.0, 1e-3), (150.0, 0.0, 0.0, 1 / 240),
                        (0.0, 30.0, 35.0, 1 / 240), (-40.0, 30.0, 35.0, 1 / 240)):
        ens = single_nucleus_ensemble(pos, vel, isp, dt_phys=dt, strengths=(S, C, P))
        ox, oy, _, _, fx, fy, amb = oracle_step(pos, vel, isp, dt, S, C, P)
        ens.step(1)
        assert pos_error(pos, ens.pos.cpu().numpy(), ox, oy, amb) <= POS_TOL
        assert force_error(ens.force.cpu().numpy(), fx, fy, amb) <= FORCE_TOL


def test_full_size_ensemble_properties():
    """Config C2 at full size (65,536 x Pb-208): a random sample of nuclei against the oracle,
    plus size-independent properties: pair forces cancel (Newton 3: total momentum after one
    step from rest is ~0 while containment is inactive) and rigidly rotated replicas stay
    rotated replicas."""
    from pyqmd_b200.state import NucleusEnsemble
    n_nuc = 65536
    ens = NucleusEnsemble.from_templates(((82, 126),), n_nuc, decay=False, keep_force=True)
    assert ens.pairs_per_step() == n_nuc * 208 * 207
    p0 = ens.pos.clone()
    ens.step(1)
    pos, vel = ens.pos.view(n_nuc, 208, 2), ens.vel.view(n_nuc, 208, 2)
    mom = vel.double().sum(1).norm(dim=1)
    tot = vel.double().norm(dim=2).sum(1)
    assert float((mom / tot).max()) < 1e-4
    # replica k and k + 64 share a template, rotated by 2*pi/1024
    import math
    ang = 2 * math.pi / 1024
    a, b = pos[0:64].double(), pos[64:128].double()
    rot = torch.stack((a[..., 0] * math.cos(ang) - a[..., 1] * math.sin(ang),
                       a[..., 0] * math.sin(ang) + a[..., 1] * math.cos(ang)), -1)
    assert float((rot - b).abs().max()) < 2e-5 * 10.0
    # sample against the oracle
    rng = np.random.default_rng(0)
    p0h, p1h = p0.view(n_nuc, 208, 2).cpu().numpy(), pos.cpu().numpy()
    isp = ens.is_proton.view(n_nuc, 208).cpu().numpy()
    fh = ens.force.view(n_nuc, 208, 2).cpu().numpy()
    for k in rng.integers(0, n_nuc, 48):
        ox, oy, _, _, fx, fy, amb = oracle_step(p0h[k], np.zeros((208, 2), np.float32), isp[k],
                                                ens.dt_phys)
        assert pos_error(p0h[k], p1h[k], ox, oy, amb) <= POS_TOL
        assert force_error(fh[k], fx, fy, amb) <= FORCE_TOL      # one step moves a nucleon by 1.5e-5 F only


def test_branch_census_matches_oracle_statistics():
    """The device-side branch census behind the roofline's FLOP count equals the oracle's
    BranchStats on the same nuclei (FP32 vs FP64 threshold flips aside)."""
    from pyqmd_b200.state import NucleusEnsemble, README_ISOTOPES
    ens = NucleusEnsemble.from_templates(README_ISOTOPES, 45, decay=False)
    ens.step(3)
    counts, flops_pair = ens.census()
    off, cnt = ens.offsets.cpu().numpy(), ens.count.cpu().numpy()
    pos, isp = ens.pos.cpu().numpy(), ens.is_proton.cpu().numpy()
    tot = {k: 0 for k in counts}
    flops = pairs = 0
    for k in range(ens.n_nuclei):
        sl = slice(off[k], off[k] + cnt[k])
        x, y = pos[sl, 0].astype(np.float64), pos[sl, 1].astype(np.float64)
        r = orc.force_step(x, y, np.zeros(cnt[k]), np.zeros(cnt[k]), isp[sl], 1 / 240,
                           integrate=False, want_stats=True)
        st = r["stats"].as_dict()
        for name in tot:
            tot[name] += st[name]
        flops += r["stats"].flops()
        pairs += cnt[k] * (cnt[k] - 1)
    for name in tot:
        assert abs(tot[name] - counts[name]) <= 4 + 1e-4 * tot[name], (name, tot[name], counts[name])
    assert abs(flops / pairs - flops_pair) < 1e-3


def test_step_arrays_equals_particle_list_path():
    """NuclearForces.step_arrays (numpy state, in place) == the list-of-Particle path, bit for bit."""
    from pyqmd_b200.forces import NuclearForces
    from pyqmd_b200.types import Particle, ParticleType
    rng = np.random.default_rng(8)
    n = 238
    x, y = rng.uniform(390, 410, n), rng.uniform(390, 410, n)
    vx, vy = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    isp = (rng.random(n) < 0.4).astype(np.uint8)
    ps = [Particle(float(a), float(b), ParticleType.PROTON if t else ParticleType.NEUTRON, float(c), float(d))
          for a, b, c, d, t in zip(x, y, vx, vy, isp)]
    nf = NuclearForces()
    nf.step(ps, 1 / 240, 3)
    nf.step_arrays(x, y, vx, vy, isp, 1 / 240, 3)
    assert np.array_equal(x, [p.x for p in ps]) and np.array_equal(vy, [p.vy for p in ps])
    with pytest.raises(ValueError):
        nf.step_arrays(x.astype(np.float32), y, vx, vy, isp, 1 / 240)


@pytest.mark.parametrize("decay", [False, True])
def test_host_buffer_pipeline_equals_device_resident_steps(decay):
    """pyqmd_ensemble_step_host (pinned host state -> chunked upload / compute / download lanes) gives
    what the device-resident ensemble gives: bit for bit with one chunk; with other chunkings a
    nucleus may sit at another place of its thread block, which regroups the per-warp reaction sums
    (FP32 rounding, ~1e-7), while every decay decision and all bookkeeping stay identical."""
    from pyqmd_b200.state import CODE_ISOTOPES, HostEnsembleRunner, NucleusEnsemble
    kw = dict(decay=decay, dt_decay=180825048000.0 * 0.05, seed=5)
    ref = NucleusEnsemble.from_templates(CODE_ISOTOPES, 9 * 40, **kw)
    ref.step(3)
    for chunks in (1, 7, 64):
        ens = NucleusEnsemble.from_templates(CODE_ISOTOPES, 9 * 40, **kw)
        run = HostEnsembleRunner(ens, chunks=chunks)
        for _ in range(3):
            run.step(1)
        torch.cuda.synchronize()
        # live slots only: slots vacated by an alpha / proton emission keep stale values
        off, cnt = ref.offsets.cpu(), ref.count.cpu().to(torch.int64)
        live = torch.zeros(ref.pos.shape[0], dtype=torch.bool)
        for o, c in zip(off.tolist(), cnt.tolist()):
            live[o:o + c] = True
        a_pos, b_pos = run.h_pos[live], ref.pos.cpu()[live]
        a_vel, b_vel = run.h_vel[live], ref.vel.cpu()[live]
        if chunks == 1:
            assert torch.equal(a_pos, b_pos) and torch.equal(a_vel, b_vel)
        else:
            assert torch.allclose(a_pos, b_pos, rtol=0, atol=2e-5), chunks
            assert torch.allclose(a_vel, b_vel, rtol=0, atol=5e-3), chunks
        if decay:
            assert torch.equal(run.h_zn, ref.zn.cpu()) and torch.equal(run.h_count, ref.count.cpu())
            assert torch.equal(run.h_isp[live], ref.is_proton.cpu()[live])
            assert int(ens.mode_counts.sum()) == int(ref.mode_counts.sum()) > 0


@pytest.mark.parametrize("n", [300, 513, 700, 1024, 1025, 1500])
def test_single_system_sizes_through_the_host_api(n):
    """One system of n nucleons through NuclearForces.step_arrays: <= 512 the pair kernel, 513..1024
    the one-nucleon-per-thread kernel, > 1024 the tiled cloud path -- all against the oracle."""
    from pyqmd_b200.forces import NuclearForces
    rng = np.random.default_rng(n)
    R = 2.5 * np.sqrt(n)
    r, th = R * np.sqrt(rng.random(n)), 2 * np.pi * rng.random(n)
    x32, y32 = (r * np.cos(th)).astype(np.float32), (r * np.sin(th)).astype(np.float32)
    isp = (rng.random(n) < 0.4).astype(np.uint8)
    x, y = x32.astype(np.float64), y32.astype(np.float64)
    vx, vy = np.zeros(n), np.zeros(n)
    p0 = np.stack([x32, y32], 1)
    ox, oy, _, _, fx, fy, amb = oracle_step(p0, np.zeros_like(p0), isp, 1 / 240)
    NuclearForces().step_arrays(x, y, vx, vy, isp, 1 / 240)
    got = np.stack([x, y], 1).astype(np.float32)
    assert pos_error(p0, got, ox, oy, amb) <= POS_TOL
    # force gate: from rest v' = 0.85 F dt (nuclear_forces.py:312-319), so F is read off the velocity
    f_dev = np.stack([vx, vy], 1) / (0.85 / 240)
    assert force_error(f_dev, fx, fy, amb) <= FORCE_TOL


@pytest.mark.parametrize("n", [257, 300, 384, 400, 511, 512])
def test_ring_kernel_with_three_and_four_warps_per_nucleus(n, monkeypatch):
    """The warp-local ring kernel at 3 and 4 groups per nucleus (257..512 nucleons; the automatic dispatch
    only takes it there for large launches): odd and even group counts, full and half off-diagonal blocks."""
    monkeypatch.setenv("PYQMD_ENSEMBLE_KERNEL", "ring")
    rng = np.random.default_rng(n)
    R = 2.2 * np.sqrt(n)
    r, th = R * np.sqrt(rng.random(n)), 2 * np.pi * rng.random(n)
    pos = np.stack([r * np.cos(th), r * np.sin(th)], 1).astype(np.float32)
    vel = (rng.standard_normal((n, 2)) * 0.1).astype(np.float32)
    isp = (rng.random(n) < 0.4).astype(np.uint8)
    ens = single_nucleus_ensemble(pos, vel, isp)
    ox, oy, ovx, ovy, fx, fy, amb = oracle_step(pos, vel, isp, ens.dt_phys)
    ens.step(1)
    assert pos_error(pos, ens.pos.cpu().numpy(), ox, oy, amb) <= POS_TOL
    assert force_error(ens.force.cpu().numpy(), fx, fy, amb) <= FORCE_TOL
    # three fused sub-steps == three single ones (state kept in shared memory / registers in between)
    a = single_nucleus_ensemble(pos, vel, isp)
    b = single_nucleus_ensemble(pos, vel, isp)
    a.step(3)
    for _ in range(3):
        b.step(1)
    assert torch.equal(a.pos, b.pos) and torch.equal(a.vel, b.vel)


@pytest.mark.parametrize("n", [600, 1000])
def test_opencl_shaped_entry_point_with_a_large_particle_list(n):
    """update_particles_gpu (float32 [n][4] buffers + caller-supplied centre, nuclear_forces.py:190-219)
    for 513..1024 particles: the ring kernel with 5..8 warps and the `center` argument."""
    from pyqmd_b200 import NuclearForces, Particle, ParticleType
    rng = np.random.default_rng(n)
    R = 2.5 * np.sqrt(n)
    r, th = R * np.sqrt(rng.random(n)), 2 * np.pi * rng.random(n)
    # FP32-representable absolute coordinates around the app's origin (400, 400), nuclear_sim.py:93
    x32 = (400.0 + r * np.cos(th)).astype(np.float32)
    y32 = (400.0 + r * np.sin(th)).astype(np.float32)
    isp = (rng.random(n) < 0.4).astype(np.uint8)
    ps = [Particle(float(a), float(b), ParticleType.PROTON if t else ParticleType.NEUTRON)
          for a, b, t in zip(x32, y32, isp)]
    NuclearForces().update_particles_gpu(ps, 1 / 240)
    x, y = x32.astype(np.float64), y32.astype(np.float64)
    vx, vy = np.zeros(n), np.zeros(n)
    res = orc.force_step(x, y, vx, vy, isp, 1 / 240, amb_tol=1e-4, want_forces=True)
    ok = ~res["amb"]
    got = np.array([[p.x, p.y] for p in ps], np.float64)
    ext = np.hypot(x32 - x32.mean(), y32 - y32.mean()).max()
    # the result is rounded to float32 ABSOLUTE coordinates (ulp 3e-5 at 400), like the reference's own buffers
    assert np.hypot(got[ok, 0] - x[ok], got[ok, 1] - y[ok]).max() / ext <= 2e-6
    f_dev = np.array([[p.vx, p.vy] for p in ps], np.float64) / (0.85 / 240)
    fx, fy = res["fx"], res["fy"]
    num = np.hypot(f_dev[ok, 0] - fx[ok], f_dev[ok, 1] - fy[ok])
    assert np.sqrt((num ** 2).sum() / (fx[ok] ** 2 + fy[ok] ** 2).sum()) <= 2e-5
