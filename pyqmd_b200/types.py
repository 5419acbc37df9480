"""Host-side mirror of the reference's domain types (OtsoBear/PyQMD particles.py:5-60) so that
code written against ``particles.ParticleType / DecayType / Particle / Nucleus`` keeps working
with the B200 path.  Enum values are the integer codes the CUDA kernels use.
"""
from __future__ import annotations

import math
import random
from enum import Enum

INF = float("inf")


class ParticleType(Enum):            # particles.py:5-11
    PROTON = 0
    NEUTRON = 1
    ALPHA = 2
    ELECTRON = 3
    GAMMA = 4
    POSITRON = 5


class DecayType(Enum):               # particles.py:13-21
    NONE = 0
    ALPHA = 1
    BETA_MINUS = 2
    BETA_PLUS = 3
    GAMMA = 4
    NEUTRON_EMISSION = 5
    PROTON_EMISSION = 6
    SPONTANEOUS_FISSION = 7


_LIFETIME = {ParticleType.ALPHA: 2.0, ParticleType.ELECTRON: 3.0, ParticleType.GAMMA: 1.0,
             ParticleType.POSITRON: 3.0}
_COLOR = {ParticleType.PROTON: (255, 100, 100), ParticleType.NEUTRON: (100, 100, 255),
          ParticleType.ALPHA: (255, 200, 0), ParticleType.ELECTRON: (0, 255, 255),
          ParticleType.GAMMA: (0, 255, 0), ParticleType.POSITRON: (255, 0, 255)}
_NUCLEONS = (ParticleType.PROTON, ParticleType.NEUTRON)


class Particle:
    """particles.py:23-50 -- attribute names are what the renderer and the force op read."""

    __slots__ = ("x", "y", "type", "vx", "vy", "radius", "lifetime", "age")

    def __init__(self, x, y, particle_type, vx=0, vy=0):
        self.x, self.y, self.type, self.vx, self.vy = x, y, particle_type, vx, vy
        self.radius = 2.5 if particle_type in _NUCLEONS else 1.0
        self.lifetime = _LIFETIME.get(particle_type, INF)
        self.age = 0

    def get_color(self):
        return _COLOR.get(self.type, (255, 255, 255))


class Nucleus:
    """particles.py:52-208: Z, N, centre, the nucleon list and the half-life (``stability``)."""

    SHELLS = (2, 8, 20, 28, 50, 82, 126)                       # particles.py:67

    def __init__(self, protons, neutrons, x, y, particles=None):
        self.protons, self.neutrons, self.x, self.y = protons, neutrons, x, y
        self.stability = 0.0                                   # set by the app, nuclear_sim.py:116
        if particles is None:
            self.particles = []
            self.initialize_particles()
        else:
            self.particles = list(particles)

    # -- initial layout: shell placement, particles.py:62-124 ------------------------------------
    def _place(self, shell_radius, want_proton):
        kind = ParticleType.PROTON if want_proton else ParticleType.NEUTRON
        radius = shell_radius * (0.8 + 0.2 * random.random())
        same = [(p.x, p.y) for p in self.particles if p.type == kind]
        best_angle, best_gap = 0, 0
        for _ in range(20):
            angle = random.uniform(0, 2 * math.pi)
            px = self.x + radius * math.cos(angle)
            py = self.y + radius * math.sin(angle)
            gap = INF
            for qx, qy in same:
                gap = min(gap, math.sqrt((qx - px) ** 2 + (qy - py) ** 2))
            if gap == INF or gap > best_gap:
                best_gap, best_angle = gap, angle
        self.particles.append(Particle(self.x + radius * math.cos(best_angle),
                                       self.y + radius * math.sin(best_angle), kind))

    def initialize_particles(self):
        total = self.protons + self.neutrons
        start_radius = 1.2 * (total ** (1 / 3)) * 0.7
        k = len(self.SHELLS)
        radii = [start_radius * (i + 1) / k for i in range(k)]
        placed_p = placed_n = 0
        shell = 0
        while placed_p < self.protons and placed_n < self.neutrons:
            pairs = min(self.SHELLS[min(shell, k - 1)] // 2,
                        min(self.protons - placed_p, self.neutrons - placed_n))
            for _ in range(pairs):
                self._place(radii[min(shell, k - 1)], True)
                self._place(radii[min(shell, k - 1)], False)
            placed_p += pairs
            placed_n += pairs
            shell = min(shell + 1, k - 1)
        for _ in range(self.protons - placed_p):
            self._place(radii[min(shell, k - 1)], True)
        for _ in range(self.neutrons - placed_n):
            self._place(radii[min(shell, k - 1)], False)

    # -- decay test, particles.py:126-147 -------------------------------------------------------------
    def should_decay(self, dt):
        from .nuclides import decay_probability
        p = decay_probability(self.stability, dt)
        if p < 0.0:
            return False                                       # stable: no draw
        return random.random() < p

    # -- transmutation of the nucleon list, particles.py:149-203 -------------------------------------
    def adjust_particles(self, decay_type):
        if decay_type in (DecayType.BETA_MINUS, DecayType.BETA_PLUS):
            src, dst = ((ParticleType.NEUTRON, ParticleType.PROTON)
                        if decay_type == DecayType.BETA_MINUS
                        else (ParticleType.PROTON, ParticleType.NEUTRON))
            for p in self.particles:
                if p.type == src:
                    p.type = dst
                    break
            return
        quota = {DecayType.ALPHA: (2, 2), DecayType.NEUTRON_EMISSION: (0, 1),
                 DecayType.PROTON_EMISSION: (1, 0)}.get(decay_type)
        if quota is None:
            return
        left = {ParticleType.PROTON: quota[0], ParticleType.NEUTRON: quota[1]}
        kept = []
        for p in self.particles:
            if left.get(p.type, 0) > 0:
                left[p.type] -= 1
            else:
                kept.append(p)
        self.particles[:] = kept
        for p in self.particles:
            p.vx *= 0.8
            p.vy *= 0.8

    def update_center_of_mass(self):                          # particles.py:205-208
        if self.particles:
            self.x = sum(p.x for p in self.particles) / len(self.particles)
            self.y = sum(p.y for p in self.particles) / len(self.particles)
