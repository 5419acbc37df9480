"""GPU tests of the headless frame driver (pyqmd_b200.sim.HeadlessSimulation): the caller side of
the hot path, nuclear_sim.py:118-176 for many nuclei at once."""
import math

import numpy as np
import pytest
import torch

from pyqmd_b200 import nuclides
from pyqmd_b200.sim import TIME_SCALE_PRESETS, HeadlessSimulation, substep_plan
from pyqmd_b200.types import DecayType, ParticleType

pytestmark = pytest.mark.gpu


def test_set_dt_decay_recomputes_probabilities_bit_exact():
    sim = HeadlessSimulation(isotopes=((6, 8), (92, 146), (82, 126), (47, 61)), n_nuclei=64)
    ens = sim.ensemble
    for dt in (1 / 240, 180825048000.0 * 1e-3, 1.409993568e17 * 0.1, 3.0):
        ens.set_dt_decay(dt)
        T = ens.half_life.cpu().numpy()
        p = ens.p_decay.cpu().numpy()
        want = np.array([nuclides.decay_probability(float(t), dt) for t in T])
        assert np.array_equal(p, want)
        tab = ens.table.cpu().numpy()
        assert np.array_equal(tab, nuclides.build_device_table(dt).view(np.uint8))


def test_frames_decay_counts_and_free_particles():
    n = 2048
    sim = HeadlessSimulation((6, 8), n_nuclei=n, seed=11)
    T = nuclides.get_half_life(6, 8)
    # pick the time scale so that a frame of 1/60 s covers T/50: 20 sub-steps of T/1000 (linear branch)
    sim.time_scale = T / 50 * 60
    frames = 12
    for _ in range(frames):
        sim.update_simulation(1 / 60)
    num_steps, eff, step_time, _ = substep_plan(1 / 60, sim.time_scale)
    assert sim.substeps_used == num_steps == 20 and eff == 1 / 240
    assert sim.ensemble.dt_phys == eff and sim.ensemble.dt_decay == step_time
    assert sim.time_passed == pytest.approx(frames / 60 * sim.time_scale)
    # C-14 -> N-14 (stable): every nucleus decays at most once, by beta-minus
    counts = sim.decay_counts
    total = sum(counts.values())
    assert counts["BETA_MINUS"] == total > 0
    zn = sim.ensemble.zn.cpu().numpy()
    assert int((zn == nuclides.zn_pack(7, 7)).sum()) == total
    p = nuclides.decay_probability(T, step_time)
    k = frames * num_steps
    expect = n * (1 - (1 - p) ** k)
    sigma = math.sqrt(n * (1 - p) ** k * (1 - (1 - p) ** k))
    assert abs(total - expect) < 5 * sigma + 1, (total, expect, sigma)
    # emitted electrons: renormalised to 50 (nuclear_sim.py:305-307) and not yet expired
    f = sim.free_particles
    assert len(f["x"]) <= total
    if len(f["x"]):
        assert set(np.unique(f["type"]).tolist()) <= {ParticleType.ELECTRON.value}
        assert np.allclose(np.hypot(f["vx"], f["vy"]), 50.0, rtol=0, atol=1e-9)
        assert (f["age"] < f["lifetime"]).all()
    # the driver drained the device event log frame by frame and saw every event
    assert sim._events_seen == total and sim.events_dropped == 0
    assert int(sim.ensemble.event_count.item()) == 0


def test_alpha_chain_with_projection_keeps_nuclei_physical():
    sim = HeadlessSimulation((92, 146), n_nuclei=96, seed=5,
                             time_scale=TIME_SCALE_PRESETS["billion"] * 2000)
    for _ in range(6):
        sim.update_simulation(1 / 60)
    ens = sim.ensemble
    cnt = ens.count.cpu().numpy()
    zn = ens.zn.cpu().numpy()
    assert sim.decay_counts["ALPHA"] > 0
    assert cnt.min() >= 238 - 4 * 8 and cnt.max() <= 238
    assert torch.isfinite(ens.pos).all()
    # render bridge: a Nucleus whose particle list mirrors the device state
    k = int(np.argmin(cnt))
    nuc = sim.nucleus_view(k)
    assert len(nuc.particles) == cnt[k]
    assert (nuc.protons, nuc.neutrons) == nuclides.zn_unpack(int(zn[k]))
    n_p = sum(1 for q in nuc.particles if q.type == ParticleType.PROTON)
    # alpha removes 2 p + 2 n, beta-minus turns n -> p: the particle list follows (Z, N) on this chain
    assert n_p == nuc.protons and len(nuc.particles) - n_p == nuc.neutrons
    xs = np.array([q.x for q in nuc.particles]); ys = np.array([q.y for q in nuc.particles])
    assert abs(xs.mean() - 400.0) < 30 and abs(ys.mean() - 400.0) < 30      # origin (400, 400), :93
    ext = np.hypot(xs - xs.mean(), ys - ys.mean()).max()
    assert 10.0 < ext < 200.0


def test_render_bridge_feeds_the_reference_renderer_signature():
    """render_args(k) = the twelve positional arguments of Renderer.render (rendering.py:32-34); the
    objects carry every attribute the renderer reads (rendering.py:42-48, 135-246)."""
    sim = HeadlessSimulation((6, 8), n_nuclei=64, seed=2)
    T = nuclides.get_half_life(6, 8)
    sim.time_scale = T / 5 * 60                      # a fifth of a half-life per frame
    for _ in range(4):
        sim.update_simulation(1 / 60)
    assert sum(sim.decay_counts.values()) > 0
    k = int(np.nonzero(sim.ensemble.zn.cpu().numpy() == nuclides.zn_pack(7, 7))[0][0])   # a decayed one
    args = sim.render_args(k)
    assert len(args) == 12
    nucleus, particles, camera_pos, zoom, time_scale, accuracy, physics_dt, substeps, max_substeps, \
        gpu_available, decay_counts, time_passed = args
    assert (nucleus.protons, nucleus.neutrons) == (7, 7) and len(nucleus.particles) == 14
    for p in sorted(nucleus.particles, key=lambda q: q.y):          # what render() does first
        assert p.radius == 2.5 and p.type in (ParticleType.PROTON, ParticleType.NEUTRON)
        assert len(p.get_color()) == 3 and abs(p.x - 400) < 200
    assert len(particles) == 1 and particles[0].type == ParticleType.ELECTRON
    fade = particles[0].age / particles[0].lifetime                 # rendering.py:47
    assert 0.0 <= fade < 1.0
    assert gpu_available is True and substeps == sim.substeps_used and max_substeps == 20
    assert set(decay_counts) == {d.name for d in DecayType if d != DecayType.NONE}
    assert time_passed == sim.time_passed and time_scale == sim.time_scale


# ---- emitted-particle life cycle on the device (pyqmd_free_particles_frame) ---------------------------
def _free_records(rows):
    from pyqmd_b200 import _lib
    rec = np.zeros(len(rows), _lib.FREE_DTYPE)
    for k, r in enumerate(rows):
        for key, v in r.items():
            rec[k][key] = v
    return rec


def test_free_particle_animation_matches_reference_goldens():
    """update_particle (nuclear_sim.py:178-210) on the device against goldens generated from the
    reference (tests/golden/sim_driver.json.gz, `animation`): six updates of one particle per case."""
    from conftest import load_json
    from pyqmd_b200.sim import FreeParticlePool, frame_constants
    gold = load_json("sim_driver.json.gz")["animation"]
    fh = float.fromhex
    pool = FreeParticlePool("cuda", capacity=64)
    for row in gold:
        ts, sub, ptype = fh(row["time_scale"]), row["substeps"], row["ptype"]
        life = 0.05 if ptype != ParticleType.NEUTRON.value else float("inf")
        p = dict(x=1.0, y=-2.0, vx=30.0, vy=-40.0, age=0.0, lifetime=life, nucleus=7, type=ptype)
        alive = []
        pool.load(_free_records([p]))
        for k in range(6):
            # one update_particle(p, 1/240, 0.004 * (k + 1)) with substeps_used = sub: a "frame" of 1 update
            f = frame_constants(ts, sub, 1 / 240, 0.004 * (k + 1), 1 / 240)
            f.num_steps = 1
            pool.frame(None, f)
            rec = pool.download()
            alive.append(len(rec) == 1)
            if len(rec) == 1:
                last = rec[0].copy()
            else:
                break
        # the reference keeps updating the object after it expired; compare up to the first expiry
        first_dead = row["alive"].index(False) if False in row["alive"] else None
        assert alive == row["alive"][: len(alive)], row
        if first_dead is None:
            assert (last["x"], last["y"], last["age"]) == (fh(row["x"]), fh(row["y"]), fh(row["age"])), row


def test_free_particle_pool_follows_the_host_mirror_over_frames():
    """A C-14 ensemble decaying over many frames: the device pool (speed / lifetime rewrite of
    handle_decay :295-342, per-sub-step animation and expiry :178-210, births mid-frame) equals the
    host mirror of pyqmd_b200.sim -- which the CPU suite pins bit for bit to the reference goldens --
    applied to the same device event log."""
    from pyqmd_b200 import sim as S
    n = 4096
    hs = HeadlessSimulation((6, 8), n_nuclei=n, seed=3)
    T = nuclides.get_half_life(6, 8)
    hs.time_scale = 0.9                              # slow motion: default lifetimes, particles expire
    host = {k: np.zeros(0) for k in ("x", "y", "vx", "vy", "age", "lifetime")}
    host["type"], host["nucleus"] = np.zeros(0, np.int32), np.zeros(0, np.int64)

    def host_animate(f, n_updates, eff, st, num_steps, ts):
        n_updates = np.broadcast_to(np.asarray(n_updates), f["x"].shape)
        for k in range(int(n_updates.max()) if len(f["x"]) else 0):
            x, y, age, alive = S.animate(f["type"], f["x"], f["y"], f["vx"], f["vy"], f["age"], f["lifetime"],
                                         eff, st, ts, num_steps)
            todo = n_updates > k
            f["x"], f["y"], f["age"] = np.where(todo, x, f["x"]), np.where(todo, y, f["y"]), np.where(todo, age, f["age"])
            keep = alive | ~todo
            f = {key: v[keep] for key, v in f.items()}
            n_updates = n_updates[keep]
        return f

    total = 0
    for frame in range(80):
        # make decays frequent without leaving slow motion: shorten the half-life the ensemble sees
        # (6 % of the C-14 decay per frame, spread over the run; electrons live 20 s = ~63 frames here)
        hs.ensemble.half_life.fill_(5.0)
        hs.ensemble.dt_decay = -1.0                  # force set_dt_decay to recompute p for the new T
        num_steps, eff, st, pdt = substep_plan(0.5, hs.time_scale)
        ens = hs.ensemble
        step0 = ens.step_index
        # replicate update_simulation, but look at the event log before the pool drains it
        ens.dt_phys = eff
        ens.set_dt_decay(st)
        ens.step(num_steps)
        ens.resolve_overlaps()
        ev = ens.events()
        ev = ev[ev["ptype"] >= 0]
        hs.substeps_used = num_steps
        hs.pool.frame(ens, S.frame_constants(hs.time_scale, num_steps, eff, st, pdt, step0))
        host = host_animate(host, num_steps, eff, st, num_steps, hs.time_scale)
        if len(ev):
            vx, vy, life = S.cosmetic_speed_lifetime_array(ev["ptype"], ev["vx"], ev["vy"], hs.time_scale,
                                                           num_steps, pdt)
            born = dict(x=ev["x"].astype(np.float64), y=ev["y"].astype(np.float64), vx=vx, vy=vy,
                        age=np.zeros(len(ev)), lifetime=life, type=ev["ptype"].astype(np.int32),
                        nucleus=ev["nucleus"].astype(np.int64))
            remaining = np.clip(num_steps - 1 - (ev["step"].astype(np.int64) - step0), 0, num_steps)
            born = host_animate(born, remaining, eff, st, num_steps, hs.time_scale)
            host = {k: np.concatenate([host[k], born[k]]) for k in host}
            total += len(ev)
    assert total > 500 and int(hs.ensemble.event_count.item()) == 0
    dev = hs.pool.download()
    order = np.lexsort((host["x"], host["type"], host["nucleus"]))
    assert len(dev) == len(order) and 0 < len(dev) < total          # some expired, some alive
    for key in ("age", "lifetime"):
        assert np.array_equal(dev[key], host[key][order]), key
    # the reference squares with ``**`` (libm pow, < 1 ulp but not correctly rounded), the device with a
    # product: |v| may differ in the last bit, and with it the renormalised velocity
    for key in ("x", "y", "vx", "vy"):
        assert np.allclose(dev[key], host[key][order], rtol=1e-15, atol=0), key
    assert np.array_equal(dev["nucleus"], host["nucleus"][order])
    assert hs.events_dropped == 0


def test_spawned_particles_get_the_reference_speed_and_lifetime():
    """handle_decay's rewrite of a fresh product (nuclear_sim.py:295-342) on the device, against the
    goldens generated from the reference (`cosmetics`: lifetime per (time scale, sub-steps, physics dt,
    particle type), speed renormalised to 30 / 50 / 60 / 40): events are written into a device event log
    by hand and absorbed by pyqmd_free_particles_frame."""
    import types as pytypes

    from conftest import load_json
    from pyqmd_b200 import _lib
    from pyqmd_b200.sim import FreeParticlePool, frame_constants
    gold = load_json("sim_driver.json.gz")["cosmetics"]
    fh = float.fromhex
    groups = {}
    for row in gold:
        groups.setdefault((row["time_scale"], row["substeps"], row["physics_dt"]), []).append(row)
    speeds = {ParticleType.ALPHA.value: 30.0, ParticleType.GAMMA.value: 60.0,
              ParticleType.ELECTRON.value: 50.0, ParticleType.POSITRON.value: 50.0}
    checked = 0
    for (ts, sub, pdt), rows in groups.items():
        ev = np.zeros(len(rows), _lib.EVENT_DTYPE)
        for k, row in enumerate(rows):
            ev[k]["nucleus"], ev[k]["step"], ev[k]["ptype"] = k, sub - 1, row["ptype"]      # born in the last sub-step
            ev[k]["x"], ev[k]["y"] = 400.0 + k, 400.0
            ev[k]["vx"], ev[k]["vy"] = 3.0 * fh(row["vx"]), 3.0 * fh(row["vy"])              # same direction, other speed
        buf = torch.from_numpy(ev.view(np.uint8).reshape(-1).copy()).cuda()
        fake = pytypes.SimpleNamespace(events_buf=buf, event_capacity=len(rows),
                                       event_count=torch.tensor([len(rows)], dtype=torch.int64, device="cuda"))
        pool = FreeParticlePool("cuda", capacity=256)
        pool.frame(fake, frame_constants(fh(ts), sub, 1 / 240, 0.004, fh(pdt), step0=0))
        rec = pool.download()
        assert len(rec) == len(rows) and int(fake.event_count.item()) == 0
        for row, r in zip(rows, rec[np.argsort(rec["nucleus"])]):
            assert r["lifetime"] == fh(row["lifetime"]), row
            assert abs(np.hypot(r["vx"], r["vy"]) - speeds.get(row["ptype"], 40.0)) < 1e-9
            assert np.allclose([r["vx"], r["vy"]], [fh(row["vx"]), fh(row["vy"])], rtol=1e-12)
            assert r["age"] == 0.0 and r["x"] == 400.0 + r["nucleus"]
            checked += 1
    assert checked == len(gold) and checked > 100
