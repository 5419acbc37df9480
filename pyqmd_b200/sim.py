"""Headless, GPU-resident counterpart of the physics half of ``NuclearSimulation``
(OtsoBear/PyQMD nuclear_sim.py:31-176, 178-210, 212-353) for N nuclei at once: the caller side
of the hot path, without pygame.

    sim = HeadlessSimulation((92, 146), n_nuclei=4096)
    sim.time_scale = 3.15576e16                 # 'billion' preset, nuclear_sim.py:86
    for _ in range(600):
        sim.update_simulation(1 / 60)           # one frame: sub-steps + overlap projection
    sim.decay_counts                            # populated (the reference never increments it)
    sim.free_particles                          # emitted alpha/e-/e+/gamma still alive

Frame logic:
  sub-step plan                    nuclear_sim.py:123-153   -> ``substep_plan`` (host)
  sub-step loop                    :161-173                 -> NucleusEnsemble.step (device)
  overlap projection once a frame  :175-176, :355-379       -> NucleusEnsemble.resolve_overlaps (device)
  emitted particle speed/lifetime  :295-342                 -> ``FreeParticlePool`` (device,
  free particle animation          :178-210                    pyqmd_free_particles_frame)
A frame issues kernels only: nothing is read back unless the caller asks for ``decay_counts`` or
``free_particles``.  ``cosmetic_speed_lifetime`` / ``animate`` are the host mirrors of the two device
kernels; the CPU test-suite pins them bit for bit to goldens generated from the reference, the GPU
suite compares the device pool with them.
"""
from __future__ import annotations

import math

import numpy as np

from . import _lib
from .types import DecayType, ParticleType

_ANIMATED = (ParticleType.ALPHA.value, ParticleType.ELECTRON.value, ParticleType.GAMMA.value,
             ParticleType.POSITRON.value)
_DEFAULT_LIFETIME = {ParticleType.ALPHA.value: 2.0, ParticleType.ELECTRON.value: 3.0,
                     ParticleType.GAMMA.value: 1.0, ParticleType.POSITRON.value: 3.0}  # particles.py:31-38

#: nuclear_sim.py:78-87
TIME_SCALE_PRESETS = {
    "real": 1.0, "minute": 60.0, "hour": 3600.0, "day": 86400.0, "year": 31557600.0,
    "millennium": 31557600000.0, "million": 31557600000000.0, "billion": 31557600000000000.0,
}


def substep_plan(dt, time_scale, physics_dt=1.0 / 240.0, accuracy=1, max_substeps=20,
                 auto_adjust_substeps=False, physics_dt_factor=0.8):
    """Sub-step plan of update_simulation (nuclear_sim.py:123-153).

    Returns (num_steps, effective_physics_dt, step_time, physics_dt) where ``step_time`` is the dt
    should_decay sees (:165) and ``effective_physics_dt`` the dt of the force step (:171)."""
    desired_dt = dt * time_scale                                            # :123
    if auto_adjust_substeps and time_scale != 1.0:                          # :131-142
        if time_scale > 1.0:
            scale = min(10.0, time_scale ** 0.3)
            physics_dt = min(1.0 / 60.0, physics_dt_factor * scale / 240.0)
        else:
            scale = max(0.1, time_scale ** 0.2)
            physics_dt = max(1.0 / 1000.0, physics_dt_factor * scale / 240.0)
    effective = physics_dt * (2.0 - accuracy)                               # :145
    factor = 1.0 if time_scale <= 10.0 else math.log10(time_scale)          # :149
    cap = int(max_substeps * factor) if auto_adjust_substeps else max_substeps   # :150
    num_steps = max(1, min(cap, int(desired_dt / effective)))               # :153
    return num_steps, effective, desired_dt / num_steps, physics_dt


def cosmetic_speed_lifetime(ptype, vx, vy, time_scale, substeps_used, physics_dt):
    """Speed renormalisation and lifetime of a freshly emitted particle (nuclear_sim.py:295-342).
    ``ptype`` is a ParticleType value; returns (vx, vy, lifetime)."""
    if ptype == ParticleType.ALPHA.value:
        base_speed = 30.0
    elif ptype == ParticleType.GAMMA.value:
        base_speed = 60.0
    elif ptype in (ParticleType.ELECTRON.value, ParticleType.POSITRON.value):
        base_speed = 50.0
    else:
        base_speed = 40.0
    mag = math.sqrt(vx ** 2 + vy ** 2)
    if mag > 0.001:
        vx, vy = (vx / mag) * base_speed, (vy / mag) * base_speed
    base_lifetime = 5.0
    if time_scale > 1.0:
        ts_factor = max(1.0, time_scale / 100.0)
        sub_factor = max(1.0, math.sqrt(substeps_used))
        dt_factor = max(1.0, 0.016 / physics_dt)
        lifetime = max(base_lifetime * sub_factor, base_lifetime * (ts_factor * sub_factor * dt_factor))
        if substeps_used > 15:
            lifetime *= substeps_used / 15.0
    else:
        default = _DEFAULT_LIFETIME.get(ptype, float("inf"))
        lifetime = max(default, base_lifetime * max(1.0, substeps_used / 5.0))
    return vx, vy, lifetime


def cosmetic_speed_lifetime_array(ptype, vx, vy, time_scale, substeps_used, physics_dt):
    """Vectorised ``cosmetic_speed_lifetime`` (same IEEE operations in the same order, so the
    results are bit-identical) for the thousands of events a large ensemble emits per frame."""
    ptype = np.asarray(ptype)
    vx, vy = np.asarray(vx, np.float64).copy(), np.asarray(vy, np.float64).copy()
    base_speed = np.full(ptype.shape, 40.0)
    base_speed[ptype == ParticleType.ALPHA.value] = 30.0
    base_speed[ptype == ParticleType.GAMMA.value] = 60.0
    base_speed[(ptype == ParticleType.ELECTRON.value) | (ptype == ParticleType.POSITRON.value)] = 50.0
    # the reference squares with ``**`` (libm pow, not guaranteed to round like a product): keep that
    # one operation in Python so every bit matches; sqrt / divide / multiply are IEEE in numpy too
    mag = np.sqrt(np.array([a ** 2 + b ** 2 for a, b in zip(vx.tolist(), vy.tolist())], np.float64))
    ok = mag > 0.001
    safe = np.where(ok, mag, 1.0)
    vx = np.where(ok, (vx / safe) * base_speed, vx)
    vy = np.where(ok, (vy / safe) * base_speed, vy)
    if time_scale > 1.0:
        _, _, life = cosmetic_speed_lifetime(ParticleType.ALPHA.value, 1.0, 0.0, time_scale,
                                             substeps_used, physics_dt)
        lifetime = np.full(ptype.shape, life)
    else:
        default = np.full(ptype.shape, float("inf"))
        for k, v in _DEFAULT_LIFETIME.items():
            default[ptype == k] = v
        lifetime = np.maximum(default, 5.0 * max(1.0, substeps_used / 5.0))
    return vx, vy, lifetime


def animate(ptype, x, y, vx, vy, age, lifetime, dt, age_dt, time_scale, substeps_used):
    """One update_particle call (nuclear_sim.py:178-210) on arrays; returns (x, y, age, alive)."""
    ptype = np.asarray(ptype)
    animated = np.isin(ptype, _ANIMATED)
    speed_scale = 0.3 * (10.0 / max(1.0, substeps_used))
    aging = min(1.0, 1.0 / (math.sqrt(max(1.0, time_scale / 100.0)) *
                            math.sqrt(max(1.0, substeps_used / 10.0))))
    step_n = dt * (time_scale ** 0.5)
    # same association as the reference: (v * ANIMATION_DT) * SPEED_SCALE  (:194-195)
    x = np.where(animated, x + vx * (1.0 / 240.0) * speed_scale, x + vx * step_n)
    y = np.where(animated, y + vy * (1.0 / 240.0) * speed_scale, y + vy * step_n)
    age = np.where(animated, age + age_dt * aging, age + age_dt)
    alive = np.where(animated, age < lifetime, True)
    return x, y, age, alive


def frame_constants(time_scale, num_steps, eff_dt, step_time, physics_dt, step0=0):
    """Per-frame constants of the free-particle update (pyqmd_free_frame), evaluated with the
    reference's expressions (nuclear_sim.py:189-190,199-200,207,320-341) in Python float64."""
    f = _lib.FreeFrame()
    f.num_steps, f.step0 = int(num_steps), int(step0) & 0xFFFFFFFF
    f.fast_forward = 1 if time_scale > 1.0 else 0
    f.speed_scale = 0.3 * (10.0 / max(1.0, num_steps))
    f.aging_scale = min(1.0, 1.0 / (math.sqrt(max(1.0, time_scale / 100.0)) *
                                    math.sqrt(max(1.0, num_steps / 10.0))))
    f.age_dt = step_time
    f.nucleon_dt = eff_dt * (time_scale ** 0.5)
    f.lifetime_fast = 0.0
    if time_scale > 1.0:            # the same for every particle type (:320-338)
        f.lifetime_fast = cosmetic_speed_lifetime(ParticleType.ALPHA.value, 1.0, 0.0, time_scale, num_steps,
                                                  physics_dt)[2]
    f.lifetime_floor = 5.0 * max(1.0, num_steps / 5.0)
    return f


class FreeParticlePool:
    """The app's ``self.particles`` (nuclear_sim.py:349) for a whole ensemble, on the device: two
    ping-pong arrays of pyqmd_free_particle and device-side counters.  ``frame`` advances the pool by
    one app frame and absorbs the ensemble's decay events of that frame; it never synchronises."""

    def __init__(self, device, capacity=1 << 20):
        from .state import resolve_device
        import torch
        self.device, self.capacity = resolve_device(device), int(capacity)
        nbytes = self.capacity * _lib.FREE_DTYPE.itemsize
        self.buf = [torch.zeros(nbytes, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.count = [torch.zeros(1, dtype=torch.int64, device=self.device) for _ in range(2)]
        self.dropped = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.cur = 0

    def frame(self, ens, frame: "_lib.FreeFrame", reset_event_count=True):
        import ctypes as C

        import torch
        a, b = self.cur, self.cur ^ 1
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().pyqmd_free_particles_frame(
                self.buf[a].data_ptr(), self.count[a].data_ptr(), self.buf[b].data_ptr(),
                self.count[b].data_ptr(), self.capacity, _lib.ptr(ens.events_buf) if ens is not None else None,
                _lib.ptr(ens.event_count) if ens is not None else None,
                ens.event_capacity if ens is not None else 0, C.byref(frame), self.dropped.data_ptr(),
                1 if reset_event_count else 0, _lib.current_stream()), "pyqmd_free_particles_frame")
        self.cur = b

    def load(self, records):
        """Replace the pool by ``records`` (structured numpy array, _lib.FREE_DTYPE)."""
        import torch
        rec = np.ascontiguousarray(records, dtype=_lib.FREE_DTYPE)
        assert len(rec) <= self.capacity
        raw = torch.from_numpy(rec.view(np.uint8).reshape(-1).copy())
        self.buf[self.cur][: raw.numel()].copy_(raw)
        self.count[self.cur].fill_(len(rec))

    def download(self):
        """The live particles as a structured numpy array sorted by (nucleus, type, x) -- the device
        order is unspecified."""
        n = min(int(self.count[self.cur].item()), self.capacity)
        raw = self.buf[self.cur][: n * _lib.FREE_DTYPE.itemsize].cpu().numpy()
        rec = raw.view(_lib.FREE_DTYPE).copy()
        return rec[np.lexsort((rec["x"], rec["type"], rec["nucleus"]))]


class HeadlessSimulation:
    """N copies of the reference's single-nucleus simulation, stepped frame by frame on the GPU."""

    def __init__(self, isotope=(92, 146), n_nuclei=1, *, isotopes=None, device="cuda", seed=0,
                 time_scale=1.0, origin=(400.0, 400.0), rotate=True, free_capacity=1 << 20):
        from .state import NucleusEnsemble
        self.isotopes = tuple(isotopes) if isotopes else (tuple(isotope),)
        self.physics_dt = 1.0 / 240.0           # nuclear_sim.py:59
        self.accuracy = 1                       # :62
        self.max_substeps = 20                  # :63
        self.auto_adjust_substeps = False       # :65
        self.physics_dt_factor = 0.8            # :66
        self.time_scale = float(time_scale)     # :50
        self.time_passed = 0.0                  # :54
        self.substeps_used = 0                  # :64
        self.frames = 0
        org = np.tile(np.asarray(origin, np.float64), (n_nuclei, 1))                # :93
        self.ensemble = NucleusEnsemble.from_templates(self.isotopes, n_nuclei, device=device,
                                                       seed=seed, origin=org, rotate=rotate,
                                                       dt_decay=1.0 / 240.0)
        # free (emitted) particles live on the device (nuclear_sim.py:349 keeps them in a list)
        self.pool = FreeParticlePool(self.ensemble.device, free_capacity)

    # -- lazily read-back views (the only places that synchronise) --------------------------------------
    @property
    def decay_counts(self):
        """Decays so far by DecayType name (:56; the reference initialises and renders this dict but
        never increments it) -- read from the device counters on demand."""
        mc = self.ensemble.mode_counts.cpu().tolist()
        return {d.name: int(mc[d.value]) for d in DecayType if d != DecayType.NONE}

    @property
    def events_dropped(self):
        """Emitted particles lost because the pool overflowed (the event log is drained every frame;
        events beyond its capacity within ONE frame are counted in decay_counts but get no particle)."""
        return int(self.pool.dropped.item())

    @property
    def _events_seen(self):
        return int(self.ensemble.mode_counts.sum().item())

    @property
    def free_particles(self):
        """The emitted particles still alive, SoA on the host (downloaded on demand)."""
        rec = self.pool.download()
        out = {k: rec[k].astype(np.float64) for k in ("x", "y", "vx", "vy", "age", "lifetime")}
        out["type"] = rec["type"].astype(np.int32)
        out["nucleus"] = rec["nucleus"].astype(np.int64)
        return out

    @property
    def free(self):
        return self.free_particles

    def update_simulation(self, dt):
        """One frame (nuclear_sim.py:118-176): kernels only, no read-back."""
        num_steps, eff_dt, step_time, self.physics_dt = substep_plan(
            dt, self.time_scale, self.physics_dt, self.accuracy, self.max_substeps,
            self.auto_adjust_substeps, self.physics_dt_factor)
        self.time_passed += dt * self.time_scale                            # :124
        self.substeps_used = num_steps                                      # :154
        ens = self.ensemble
        ens.dt_phys = eff_dt
        ens.set_dt_decay(step_time)
        step0 = ens.step_index
        ens.step(num_steps)                                                 # :161-173
        ens.resolve_overlaps()                                              # :175-176
        # free particles: those already alive get one update per sub-step (:162); this frame's emissions
        # (:349) get the speed / lifetime rewrite (:295-342) and the sub-steps left after their own
        self.pool.frame(ens, frame_constants(self.time_scale, num_steps, eff_dt, step_time, self.physics_dt,
                                             step0))
        self.frames += 1

    def free_particle_objects(self, k=None, types=None):
        """The emitted particles still alive as ``Particle`` objects (what the app keeps in
        ``self.particles``, nuclear_sim.py:349); ``k``: only those emitted by nucleus ``k``.
        ``types``: the module providing ``Particle`` / ``ParticleType`` (default: pyqmd_b200.types; pass
        the reference's ``particles`` module to hand its own classes to its own ``Renderer``, which
        compares ``particle.type`` with ITS enum, rendering.py:72,81)."""
        from . import types as own
        T = types or own
        f = self.free_particles
        sel = np.arange(len(f["x"])) if k is None else np.nonzero(f["nucleus"] == k)[0]
        out = []
        for i in sel:
            p = T.Particle(float(f["x"][i]), float(f["y"][i]), T.ParticleType(int(f["type"][i])),
                           float(f["vx"][i]), float(f["vy"][i]))
            p.age, p.lifetime = float(f["age"][i]), float(f["lifetime"][i])
            out.append(p)
        return out

    def render_args(self, k=0, camera_pos=(400.0, 400.0), zoom=15.0, types=None):
        """Positional arguments of the reference's ``Renderer.render`` (rendering.py:32-34, called at
        nuclear_sim.py:598-603) for nucleus ``k``: pygame can draw the GPU-resident state with
        ``renderer.render(*sim.render_args(k, types=particles))``."""
        return (self.nucleus_view(k, types), self.free_particle_objects(k, types), list(camera_pos), zoom,
                self.time_scale, self.accuracy, self.physics_dt, self.substeps_used, self.max_substeps,
                True, dict(self.decay_counts), self.time_passed)

    def nucleus_view(self, k=0, types=None):
        """A ``Nucleus`` materialised from the device state of nucleus ``k`` -- the render bridge:
        ``Renderer`` reads ``.particles[i].x/.y/.type/.radius`` and ``.protons/.neutrons/.stability``
        (rendering.py:42-48, 135-246).  ``types``: see ``free_particle_objects``."""
        from . import types as own
        T = types or own
        ens = self.ensemble
        o, c = int(ens.offsets[k]), int(ens.count[k])
        pos = ens.pos[o:o + c].cpu().numpy().astype(np.float64)
        vel = ens.vel[o:o + c].cpu().numpy().astype(np.float64)
        isp = ens.is_proton[o:o + c].cpu().numpy()
        org = ens.origin[k].cpu().numpy() if ens.origin is not None else np.zeros(2)
        ps = [T.Particle(float(p[0] + org[0]), float(p[1] + org[1]),
                         T.ParticleType.PROTON if t else T.ParticleType.NEUTRON, float(v[0]), float(v[1]))
              for p, v, t in zip(pos, vel, isp)]
        zn = int(ens.zn[k])
        if T is own:
            nuc = own.Nucleus(zn >> 16, zn & 0xFFFF, float(org[0]), float(org[1]), particles=ps)
        else:       # a foreign Nucleus class would lay out fresh particles in its constructor (particles.py:62)
            nuc = object.__new__(T.Nucleus)
            nuc.protons, nuc.neutrons = zn >> 16, zn & 0xFFFF
            nuc.x, nuc.y = float(org[0]), float(org[1])
            nuc.particles, nuc.decay_chain = ps, []
        nuc.update_center_of_mass()
        nuc.stability = float(ens.half_life[k])
        return nuc
