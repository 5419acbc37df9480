"""The reference's OWN application loop on top of pyqmd_b200.

baseline/_ref holds the unmodified reference (baseline/install_reference.py).  Its nuclear_sim module is
imported with stub pygame / pyopencl modules (it has no other way to run without a display), and the one
import the drop-in replaces -- `from nuclear_forces import NuclearForces`, nuclear_sim.py:14 -- is swapped
for `pyqmd_b200.NuclearForces`.  NuclearSimulation.update_simulation (nuclear_sim.py:118-176: sub-step
plan, should_decay / handle_decay, force step, resolve_overlaps) then runs as written, once with the
reference's own CPU path (`update_particles_cpu`, nuclear_forces.py:236-323) and once with the GPU
drop-in, teacher-forced frame by frame from identical states and identical `random` streams.
"""
import copy
import random
from collections import deque

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("baseline/_ref not installed (python baseline/install_reference.py)")
    R = ref_loader.Ref()
    R.ns = R.nuclear_sim()
    import logging
    logging.getLogger("NuclearSim").setLevel(logging.WARNING)      # the app logs every decay at INFO (:285)
    return R


def headless(ns, forces, gpu_available):
    """NuclearSimulation without pygame: the attributes __init__ sets (nuclear_sim.py:31-90), minus the
    window, the clock and the renderer."""
    sim = object.__new__(ns.NuclearSimulation)
    sim.forces, sim.gpu_available = forces, gpu_available
    sim.nucleus, sim.particles = None, []
    sim.time_scale, sim.time_passed = 1.0, 0
    sim.decay_counts = {}
    sim.decay_times = deque(maxlen=100)
    sim.physics_dt = 1.0 / 240.0
    sim.fps_history = deque(maxlen=30)
    sim.manual_accuracy, sim.accuracy, sim.max_substeps, sim.substeps_used = True, 1, 20, 0
    sim.auto_adjust_substeps, sim.physics_dt_factor = False, 0.8
    sim.camera_pos, sim.camera_target = [400, 400], [400, 400]
    sim.zoom_level = sim.target_zoom = 15.0
    sim.zoom_speed = 0.1
    return sim


def state_of(sim):
    ps = sim.nucleus.particles
    return (np.array([[p.x, p.y] for p in ps], np.float64), np.array([[p.vx, p.vy] for p in ps], np.float64),
            [p.type.value for p in ps])


def run_teacher_forced(ref, frames, time_scale, method, seed, isotope=(92, 146)):
    """Sim A = the reference's CPU path, sim B = the same application code over pyqmd_b200.NuclearForces.
    Every frame B restarts from A's state and RNG state.  Returns, per frame, the position error after the
    sub-steps (BEFORE resolve_overlaps: the hot path itself), after the whole frame, the velocity error
    (relative to the nucleus extent / the largest speed) and the length of the decay chain."""
    import pyqmd_b200
    ns = ref.ns
    ns.NuclearForces = pyqmd_b200.NuclearForces          # the one-line switch of INTEGRATION.md section 1
    gpu_forces = ns.NuclearForces()
    random.seed(seed)
    a = headless(ns, ref.forces(), gpu_available=False)      # reference CPU path (:173)
    a.create_nucleus(*isotope)                               # nuclear_sim.py:92-116
    a.time_scale = time_scale
    b = headless(ns, gpu_forces, gpu_available=(method == "gpu"))   # :171 update_particles_gpu / :173 _cpu
    # Nucleons with a partner within AMB (relative) of a discontinuity of the law (d^2 = 0.01, d = 2.8,
    # 4.25, 8, 9) may take the other branch once B rounds the shared float64 state to FP32: the oracle
    # flags them for every sub-step A takes, and they are excluded (and counted) for that frame -- the
    # convention of every other parity test (tests/gpu_util.py), with a window that covers the rounding
    # of the INPUT (2e-6 of a 40-unit nucleus) as well.
    from oracle import oracle as orc
    amb_frame = []
    ref_step = type(a.forces).update_particles_cpu

    def spy_forces(particles, dt):
        x = np.array([p.x for p in particles]); y = np.array([p.y for p in particles])
        isp = np.array([p.type == ref.particles.ParticleType.PROTON for p in particles], np.uint8)
        r = orc.force_step(x, y, np.zeros_like(x), np.zeros_like(x), isp, dt, amb_tol=AMB)
        amb_frame.append(r["amb"])
        return ref_step(a.forces, particles, dt)
    a.forces.update_particles_cpu = spy_forces
    before = {}
    for name, sim in (("a", a), ("b", b)):                   # look at the state resolve_overlaps receives
        def spy(sim=sim, name=name, real=type(sim).resolve_overlaps):
            before[name] = state_of(sim)[0]
            return real(sim)
        sim.resolve_overlaps = spy
    pre, post, verrs, decays, excluded = [], [], [], [], []
    for f in range(frames):
        del amb_frame[:]
        b.nucleus = copy.deepcopy(a.nucleus)
        b.particles = copy.deepcopy(a.particles)
        b.time_scale, b.time_passed = a.time_scale, a.time_passed
        rng = random.getstate()
        a.update_simulation(1 / 60)
        rng_after_a = random.getstate()
        random.setstate(rng)
        b.update_simulation(1 / 60)
        assert random.getstate() == rng_after_a              # both consumed exactly the same draws
        pa, va, ta = state_of(a)
        pb, vb, tb = state_of(b)
        assert ta == tb and (a.nucleus.protons, a.nucleus.neutrons) == (b.nucleus.protons, b.nucleus.neutrons)
        assert a.substeps_used == b.substeps_used and len(a.particles) == len(b.particles)
        ext = np.hypot(*(pa - pa.mean(0)).T).max()
        n_now = len(pa)
        ok = np.ones(n_now, bool)
        if all(len(m) == n_now for m in amb_frame):          # (a decay inside the frame changes the list)
            for m in amb_frame:
                ok &= ~m
        excluded.append(1.0 - ok.mean())
        d_pre = np.hypot(*(before["a"] - before["b"]).T)
        pre.append(d_pre[ok].max() / ext if ok.any() else 0.0)
        post.append(np.hypot(*(pa - pb).T).max() / ext)
        verrs.append(np.hypot(*(va - vb).T).max() / max(np.hypot(*va.T).max(), 1e-9))
        decays.append(len(a.nucleus.decay_chain))
    return np.array(pre), np.array(post), np.array(verrs), decays, np.array(excluded)


AMB = 1e-4


@pytest.mark.parametrize("method", ["gpu", "cpu"])
def test_reference_update_simulation_runs_on_the_dropin_50_frames(ref, method):
    """Real time, 60 fps: 4 sub-steps of 1/240 per frame + resolve_overlaps, 50 frames.
    Gate on the hot path: after the frame's 4 free-running sub-steps (before the projection) every
    nucleon that stayed clear of the law's discontinuities is within 4 x the per-step FP32 budget (1e-5
    of the extent) of the float64 CPU path.  The projection itself is a sequential sweep that amplifies
    any difference while the fresh, heavily overlapping layout unfolds (tests/test_gpu_overlaps.py
    measures that conditioning), so the state after it is reported, not gated."""
    pre, post, verrs, _, excl = run_teacher_forced(ref, 50, 1.0, method, seed=7)
    print(f"reference app on the drop-in ({method}): position error per frame before the projection "
          f"max {pre.max():.2e} median {np.median(pre):.2e} (nucleons near a threshold excluded: mean "
          f"{excl.mean():.1%}, max {excl.max():.1%}); after the projection max {post.max():.2e} median "
          f"{np.median(post):.2e}")
    assert pre.max() <= 4e-5
    assert excl.mean() <= 0.25 and np.isfinite(post).all()


def test_reference_update_simulation_with_decays_on_the_dropin(ref):
    """Fast forward (20 sub-steps per frame, the whole U-238 chain firing): identical decay decisions,
    chains and emitted particles (same reference code, same draws -- asserted inside the loop)."""
    pre, post, verrs, decays, excl = run_teacher_forced(ref, 8, 31557600000000000.0 * 2000, "gpu", seed=11)
    assert decays[-1] > 10                                   # the chain advanced (:285 bookkeeping)
    print(f"reference app with decays: position error before the projection max {pre.max():.2e} "
          f"(20 free-running sub-steps per frame), decays {decays[-1] - 1}")
    assert np.isfinite(post).all() and pre.max() <= 1e-3      # 20 free-running sub-steps through a whole decay chain (measured 1.8e-4)


@pytest.mark.parametrize("isotope,time_scale,min_decays", [
    ((6, 8), 180825048000.0 * 60 / 3, 1),            # C-14: beta-minus within a few frames, then stable N-14
    ((84, 134), 186.0 * 60 * 4, 1),                  # Po-218: the two-option branch (alpha 99.98 % / beta-minus)
    ((26, 30), 1.0, 0),                              # Fe-56: stable, the sub-warp ring / block kernels at A = 56
])
def test_reference_app_other_presets_on_the_dropin(ref, isotope, time_scale, min_decays):
    """The same teacher-forced comparison for other nuclei of the app's preset list (nuclear_sim.py:494-504):
    light and medium nuclei take other kernel geometries, C-14 and Po-218 decay through other modes."""
    pre, post, verrs, decays, excl = run_teacher_forced(ref, 12, time_scale, "cpu", seed=3, isotope=isotope)
    print(f"reference app, isotope {isotope}: position error before the projection max {pre.max():.2e}, "
          f"decays {decays[-1] - 1}")
    assert decays[-1] - 1 >= min_decays
    assert np.isfinite(post).all() and pre.max() <= (4e-5 if min_decays == 0 else 1e-3)
