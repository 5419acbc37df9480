"""Float64 CPU restatement of the reference's decay slice (tables, heuristics, transmutation,
emission, sub-step order).

TEST INFRASTRUCTURE ONLY -- see oracle/pyqmd_oracle.c.  Pinned by tests/golden/decay_*.json,
which tests/golden/gen_golden.py produced by calling the unmodified reference functions.

Draw slots: per nucleus-step the reference may consume up to four uniforms
from its global stream; here they are explicit arguments:
  slot 0  should_decay            particles.py:147          (only if half-life finite)
  slot 1  branch pick             decay_chains.py:221       (only if > 1 option)
  slot 2  emission angle          decay_chains.py:332..367  (only if a particle is emitted)
  slot 3  daughter half-life      decay_chains.py:312-328   (only if it must be estimated)
"""
from __future__ import annotations

import json
import math
import os

import numpy as np

from . import oracle as _orc

INF = float("inf")
YEAR, DAY, HOUR, MINUTE = 31557600.0, 86400.0, 3600.0, 60.0   # decay_chains.py:6-9

# DecayType values, particles.py:13-21
NONE, ALPHA, BETA_MINUS, BETA_PLUS, GAMMA, NEUTRON_EMISSION, PROTON_EMISSION, FISSION = range(8)
# ParticleType values, particles.py:5-11
PROTON, NEUTRON, P_ALPHA, ELECTRON, P_GAMMA, POSITRON = range(6)

# emitted particle type and speed per decay mode, decay_chains.py:331-371
EMISSION = {
    ALPHA: (P_ALPHA, 100), BETA_MINUS: (ELECTRON, 150), BETA_PLUS: (POSITRON, 150),
    GAMMA: (P_GAMMA, 200), NEUTRON_EMISSION: (NEUTRON, 60), PROTON_EMISSION: (PROTON, 50),
}

_DATA = None


def data():
    """HALF_LIVES (decay_chains.py:13-123) and DECAY_CHAINS (:126-167) as dumped from the
    reference at import time by gen_golden.py (values stored as float.hex strings)."""
    global _DATA
    if _DATA is None:
        with open(os.path.join(os.path.dirname(__file__), "nuclide_data.json")) as f:
            raw = json.load(f)
        hl = {(z, n): float.fromhex(h) for z, n, h in raw["half_lives"]}
        ch = {(z, n): [(a, b, m, float.fromhex(p)) for a, b, m, p in opts]
              for z, n, opts in raw["decay_chains"]}
        _DATA = (hl, ch)
    return _DATA


def stable_ratio(z, strict):
    """decay_chains.py:182-187 uses ``z < 20``; :279-282 also ``z < 20``."""
    return 1.0 if z < 20 else 1.0 + 0.015 * z ** 1.3


def decay_options(z, n):
    """Options [(Z', N', mode, prob)] for nuclide (z, n): the table entry, else the heuristic
    single option of expand_decay_chain (decay_chains.py:169-201)."""
    _, chains = data()
    if (z, n) in chains:
        return chains[(z, n)]
    n_to_z = n / max(1, z)                                   # :178
    sr = stable_ratio(z, True)                               # :182-187
    if z > 83:                                               # :190-191
        return [(z - 2, n - 2, ALPHA, 0.9)]
    if n_to_z > sr + 0.15:                                   # :192-193
        return [(z + 1, n - 1, BETA_MINUS, 0.9)]
    if n_to_z < sr - 0.15:                                   # :194-198
        if z > 30:
            return [(z - 1, n + 1, BETA_PLUS, 0.9)]
        return [(z - 1, n, PROTON_EMISSION, 0.9)]
    return [(z, n, NONE, 1.0)]                               # :201


def pick_option(options, r):
    """decay_chains.py:218-229: single option taken unconditionally (no draw); otherwise the
    first option with r <= running sum, falling through to option 0."""
    if len(options) == 1:
        return 0
    cum = 0
    for k, (_, _, _, p) in enumerate(options):
        cum += p
        if r <= cum:
            return k
    return 0


def decay_product(z, n, r=None):
    """get_decay_product (decay_chains.py:203-245) -> (Z', N', mode or None, draws_used)."""
    opts = decay_options(z, n)
    used = 0
    if len(opts) > 1:
        used = 1
    k = pick_option(opts, r)
    nz, nn, mode, _ = opts[k]
    if mode == NONE:                                         # :231-232
        return z, n, None, used
    return nz, nn, mode, used


def half_life_class(z, n):
    """get_half_life (decay_chains.py:247-328) split into its deterministic part.

    Returns ('table', T) for a database hit (:257-262; T may be inf), ('inf',) when the
    stability score is >= 0.95 (:309-310), or ('band', a, b, unit) meaning
    ``10 ** uniform(a, b) * unit`` (:311-328; unit 1.0 for the last band)."""
    hl, _ = data()
    if (z, n) in hl:
        return ("table", hl[(z, n)])
    n_to_z = n / max(1, z)                                   # :278
    sr = stable_ratio(z, False)                              # :279-282
    deviation = abs(n_to_z - sr)                             # :284
    magic = (2, 8, 20, 28, 50, 82, 126)                      # :287
    bonus = 0
    if z in magic:
        bonus += 0.2                                         # :289-290
    if n in magic:
        bonus += 0.2                                         # :291-292
    parity = 1.0                                             # :295-299
    if z % 2 == 0 and n % 2 == 0:
        parity = 0.5
    elif z % 2 == 1 and n % 2 == 1:
        parity = 2.0
    stability = max(0, 1.0 - deviation * 2.0 - parity * 0.1 + bonus)   # :302
    if z > 83:
        stability *= 0.5                                     # :305-306
    if stability >= 0.95:
        return ("inf",)
    for thr, a, b, unit in ((0.85, 15, 17, YEAR), (0.75, 9, 14, YEAR), (0.65, 6, 9, YEAR),
                            (0.50, 3, 6, YEAR), (0.40, 0, 3, YEAR), (0.30, 0, 2, DAY),
                            (0.20, 0, 4, HOUR), (0.10, -1, 3, MINUTE)):
        if stability >= thr:
            return ("band", a, b, unit)
    return ("band", -6, 1, 1.0)                              # :328 (no unit factor)


def half_life(z, n, u=None):
    """get_half_life value; ``u`` is the slot-3 uniform when the class is a band.
    Returns (T, draws_used)."""
    c = half_life_class(z, n)
    if c[0] == "table":
        return c[1], 0
    if c[0] == "inf":
        return INF, 0
    _, a, b, unit = c
    e = a + (b - a) * u                                      # random.uniform(a, b)
    if unit == 1.0 and a == -6:
        return 10 ** e, 1                                    # :328
    return 10 ** e * unit, 1


def adjust_types(types, mode):
    """Nucleus.adjust_particles (particles.py:149-203) on a list of PROTON/NEUTRON codes.

    Returns (new_types, removed_indices (ascending), damp) where ``damp`` says whether the
    survivors' velocities are scaled by 0.8 (:201-203; only on the removal path)."""
    types = list(types)
    if mode == BETA_MINUS:                                   # :158-164
        for i, t in enumerate(types):
            if t == NEUTRON:
                types[i] = PROTON
                break
        return types, [], False
    if mode == BETA_PLUS:                                    # :165-171
        for i, t in enumerate(types):
            if t == PROTON:
                types[i] = NEUTRON
                break
        return types, [], False
    if mode == ALPHA:
        rp, rn = 2, 2                                        # :155-157
    elif mode == NEUTRON_EMISSION:
        rp, rn = 0, 1                                        # :172-174
    elif mode == PROTON_EMISSION:
        rp, rn = 1, 0                                        # :175-177
    else:
        return types, [], False                              # :178-179
    removed = []
    for i, t in enumerate(types):                            # :183-192
        if rp > 0 and t == PROTON:
            removed.append(i)
            rp -= 1
        elif rn > 0 and t == NEUTRON:
            removed.append(i)
            rn -= 1
        if rp == 0 and rn == 0:
            break
    keep = [t for i, t in enumerate(types) if i not in set(removed)]
    return keep, removed, True


class OracleNucleus:
    """State of one nucleus + the reference's per-sub-step order (nuclear_sim.py:161-173 and
    the physics slice of handle_decay :213,288-294,349,353)."""

    def __init__(self, z, n, x, y, is_proton, vx=None, vy=None, origin=(0.0, 0.0), T=None):
        self.z, self.n = int(z), int(n)
        self.x = np.array(x, np.float64)
        self.y = np.array(y, np.float64)
        self.vx = np.zeros(len(self.x)) if vx is None else np.array(vx, np.float64)
        self.vy = np.zeros(len(self.x)) if vy is None else np.array(vy, np.float64)
        self.types = [PROTON if t else NEUTRON for t in is_proton]
        self.cx, self.cy = origin                  # Nucleus.x / .y (particles.py:56-57)
        if T is None:
            T, _ = half_life(self.z, self.n, 0.5)  # nuclear_sim.py:116
        self.T = T
        self.emitted = []                          # (type, x, y, vx, vy)

    def decay_event(self, u_branch, u_angle, u_half):
        """Physics slice of handle_decay.  Returns (mode or None, draws used per slot 1..3)."""
        nz, nn, mode, used1 = decay_product(self.z, self.n, u_branch)    # nuclear_sim.py:213
        if mode is None:                                                # :215
            return None, (used1, 0, 0)
        self.z, self.n = nz, nn                                         # :288-289
        new_types, removed, damp = adjust_types(self.types, mode)       # :290
        if removed:
            keep = np.ones(len(self.types), bool)
            keep[removed] = False
            self.x, self.y = self.x[keep].copy(), self.y[keep].copy()
            self.vx, self.vy = self.vx[keep].copy(), self.vy[keep].copy()
        self.types = new_types
        if damp:                                                        # particles.py:201-203
            self.vx *= 0.8
            self.vy *= 0.8
        if len(self.x):                                                 # :291, particles.py:205-208
            self.cx = _orc.py312_mean(self.x)
            self.cy = _orc.py312_mean(self.y)
        used2 = 0
        if mode in EMISSION:                                            # :294, decay_chains.py:331-371
            ptype, speed = EMISSION[mode]
            angle = 0 + (2 * math.pi - 0) * u_angle
            self.emitted.append((ptype, self.cx, self.cy, speed * math.cos(angle),
                                 speed * math.sin(angle)))
            used2 = 1
        self.T, used3 = half_life(self.z, self.n, u_half)               # :353
        return mode, (used1, used2, used3)

    def substep(self, dt_phys, dt_decay, draws, S=150.0, Cc=30.0, P=35.0):
        """One pass of the loop body nuclear_sim.py:165-173.  ``draws`` = 4 uniforms (slots).
        Returns (decayed, mode, consumed[4])."""
        consumed = [0, 0, 0, 0]
        decayed, mode = False, None
        p = _orc.decay_probability(self.T, dt_decay)                    # particles.py:126-144
        if p >= 0.0:
            consumed[0] = 1
            if draws[0] < p:                                            # :147
                decayed = True
                mode, (c1, c2, c3) = self.decay_event(draws[1], draws[2], draws[3])
                consumed[1:] = [c1, c2, c3]
        if len(self.x) > 0:                                             # nuclear_sim.py:169
            t = np.array([1 if k == PROTON else 0 for k in self.types], np.uint8)
            _orc.force_step(self.x, self.y, self.vx, self.vy, t, dt_phys, S, Cc, P)
        return decayed, mode, consumed
