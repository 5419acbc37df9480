#!/usr/bin/env python
"""Generate the golden fixtures by running the UNMODIFIED reference (OtsoBear/PyQMD).

Run in the build container only (needs /root/reference or $PYQMD_REF):

    python tests/golden/gen_golden.py

Writes
  tests/golden/force_kats.json        known-answer force steps (nuclear_forces.py:236-323)
  tests/golden/u238_traj.npz          U-238 1000-step free-running trajectory snapshots
  tests/golden/decay_tables.json.gz   get_half_life / get_decay_product over a (Z,N) grid
  tests/golden/decay_events.json.gz   should_decay probabilities, seeded decision strings,
                                      adjust_particles cases, decay-chain walks, sub-step loops
  tests/golden/resolve_overlaps.json.gz  per-frame overlap projection (nuclear_sim.py:355-379)
  tests/golden/sim_driver.json.gz     sub-step plan, emitted-particle cosmetics and animation
  oracle/nuclide_data.json            HALF_LIVES / DECAY_CHAINS dump (oracle's copy)
  pyqmd_b200/data/nuclide_data.json   same dump (product's copy)
  pyqmd_b200/data/layout_templates.npz  reference-generated initial layouts (particles.py:62-124)
                                      of the nine preset isotopes, 64 seeds each

All floats that must be reproduced bit-for-bit are stored as float.hex() strings.
The reference stays unmodified: stubs for pyopencl / pygame / siphash24 are injected and the
module-level ``random`` of particles.py / decay_chains.py is replaced by a DrawFeeder where
explicit draws are needed (oracle/ref_loader.py).
"""
import gzip
import json
import logging
import math
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
R = ref_loader.Ref()
PT, DT = R.particles.ParticleType, R.particles.DecayType
logging.getLogger("NuclearSim").setLevel(logging.ERROR)

hx = float.hex


def hexlist(v):
    return [hx(float(a)) for a in v]


# readme.md:43-51 isotope list (BASELINE config 3) and nuclear_sim.py:494-504 code list
README_ISOTOPES = [(1, 0), (2, 2), (6, 6), (6, 8), (26, 30), (47, 60), (79, 118), (82, 126),
                   (92, 146)]
CODE_ISOTOPES = [(1, 2), (2, 3), (6, 8), (8, 9), (26, 33), (47, 61), (79, 119), (82, 127),
                 (92, 146)]


def state_of(ps):
    return dict(x=hexlist(p.x for p in ps), y=hexlist(p.y for p in ps),
                vx=hexlist(p.vx for p in ps), vy=hexlist(p.vy for p in ps),
                is_proton=[int(p.type == PT.PROTON) for p in ps])


# ------------------------------------------------------------------------------------------
def gen_nuclide_data():
    dc = R.decay_chains
    out = {
        "source": "decay_chains.py:13-123 (HALF_LIVES), :126-167 (DECAY_CHAINS) at import time",
        "half_lives": [[z, n, hx(float(v))] for (z, n), v in dc.HALF_LIVES.items()],
        "decay_chains": [[z, n, [[a, b, m.value, hx(float(p))] for a, b, m, p in opts]]
                         for (z, n), opts in dc.DECAY_CHAINS.items()],
    }
    for path in (os.path.join(ROOT, "oracle", "nuclide_data.json"),
                 os.path.join(ROOT, "pyqmd_b200", "data", "nuclide_data.json")):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            json.dump(out, f, separators=(",", ":"))
    print("nuclide data:", len(out["half_lives"]), "half-lives,", len(out["decay_chains"]), "chains")


# ------------------------------------------------------------------------------------------
def gen_force_kats():
    nf = R.forces()
    cases = []

    def run(name, xs, ys, types, vxs=None, vys=None, dt=1 / 240, S=150.0, C=30.0, P=35.0,
            steps=1):
        n = len(xs)
        vxs = vxs or [0.0] * n
        vys = vys or [0.0] * n
        f = R.forces(S, C, P)
        ps = R.make_particles(xs, ys, vxs, vys, types)
        inp = state_of(ps)
        for _ in range(steps):
            f.update_particles_cpu(ps, dt)
        cases.append(dict(name=name, dt=hx(dt), S=hx(S), C=hx(C), P=hx(P), steps=steps,
                          input=inp, output=state_of(ps)))

    # SURVEY.md section 4 KATs
    run("A_pp_d3", [0, 3], [0, 0], [1, 1])
    run("B_pn_d2", [0, 0], [0, 2], [1, 0])
    run("C_nn_d10", [0, 6], [0, 8], [0, 0])
    run("C_pp_d10", [0, 6], [0, 8], [1, 1])
    run("D_pp_skip", [0, 0.05], [0, 0.05], [1, 1])
    run("E_nn_containment", [0, 100], [0, 0], [0, 0])
    run("single", [1.5], [2.5], [1], [0.25], [-0.5])
    # one case per branch boundary region, both type combinations
    for d in (0.05, 0.11, 1.0, 2.79, 2.81, 4.2, 4.3, 7.9, 8.1, 8.99, 9.01, 15.0, 40.0):
        for types in ([1, 1], [1, 0], [0, 0]):
            run(f"pair_d{d}_{types[0]}{types[1]}", [0.0, d * 0.6], [0.0, d * 0.8], types)
    # non-default strengths and dt, clamp active / inactive
    run("strengths", [0, 5, 1], [0, 1, 7], [1, 1, 0], S=20.0, C=3.0, P=4.0, dt=1 / 60)
    run("strengths_big", [0, 9.5, 1], [0, 1, 12], [1, 1, 0], S=900.0, C=300.0, P=4.0, dt=1 / 1000)
    # random systems, FP32-representable inputs, several steps
    rng = random.Random(2024)
    for n, ext, steps in ((5, 6.0, 3), (14, 5.0, 5), (33, 12.0, 3), (64, 30.0, 2),
                          (100, 8.0, 2)):
        xs = [float(np.float32(rng.uniform(-ext, ext))) for _ in range(n)]
        ys = [float(np.float32(rng.uniform(-ext, ext))) for _ in range(n)]
        vxs = [float(np.float32(rng.uniform(-1, 1))) for _ in range(n)]
        vys = [float(np.float32(rng.uniform(-1, 1))) for _ in range(n)]
        types = [int(rng.random() < 0.4) for _ in range(n)]
        run(f"random_n{n}", xs, ys, types, vxs, vys, steps=steps)
    with open(os.path.join(GOLD, "force_kats.json"), "w") as f:
        json.dump(dict(source="NuclearForces.update_particles_cpu, nuclear_forces.py:236-323",
                       cases=cases), f, separators=(",", ":"))
    print("force KATs:", len(cases))


# ------------------------------------------------------------------------------------------
def gen_u238_traj(n_steps=1000):
    """Config C1: random.seed(0); Nucleus(92,146,400,400); positions rounded to FP32; v = 0;
    dt = 1/240; free-running reference trajectory.  Snapshot pairs (state at s, state at s+1)
    are kept for teacher-forced per-step parity; every 50th state for drift reporting."""
    random.seed(0)
    nuc = R.particles.Nucleus(92, 146, 400, 400)
    ps = nuc.particles
    for p in ps:
        p.x = float(np.float32(p.x))
        p.y = float(np.float32(p.y))
    nf = R.forces()
    types = np.array([p.type == PT.PROTON for p in ps], np.uint8)
    pair_steps = [0, 1, 2, 3, 5, 10, 20, 49, 50, 100, 200, 350, 500, 750, 998, 999]
    keep = set(pair_steps) | {s + 1 for s in pair_steps} | set(range(0, n_steps + 1, 50))
    snaps = {}

    def snap(s):
        snaps[s] = np.array([[p.x, p.y, p.vx, p.vy] for p in ps], np.float64)

    t0 = time.time()
    snap(0)
    for s in range(1, n_steps + 1):
        nf.update_particles_cpu(ps, 1 / 240)
        if s in keep:
            snap(s)
    steps = np.array(sorted(snaps), np.int32)
    np.savez_compressed(os.path.join(GOLD, "u238_traj.npz"), steps=steps,
                        states=np.stack([snaps[int(s)] for s in steps]), is_proton=types,
                        pair_steps=np.array(pair_steps, np.int32), dt=np.float64(1 / 240))
    print(f"u238 trajectory: {n_steps} steps in {time.time() - t0:.1f}s, {len(steps)} snapshots")


# ------------------------------------------------------------------------------------------
def gen_layouts(n_seeds=64):
    """Initial layouts from Nucleus.initialize_particles (particles.py:62-124), origin (0,0),
    random.seed(1000*k + seed) for isotope k; stored FP32 (the device state precision)."""
    out = {}
    t0 = time.time()
    for name, iso_list in (("readme", README_ISOTOPES), ("code", CODE_ISOTOPES)):
        for k, (z, n) in enumerate(iso_list):
            key = f"z{z}_n{n}"
            if key + "_xy" in out:
                continue
            a = z + n
            xy = np.zeros((n_seeds, a, 2), np.float32)
            tp = np.zeros((n_seeds, a), np.uint8)
            for s in range(n_seeds):
                random.seed(1000 * (z * 256 + n) + s)
                nuc = R.particles.Nucleus(z, n, 0.0, 0.0)
                xy[s, :, 0] = [p.x for p in nuc.particles]
                xy[s, :, 1] = [p.y for p in nuc.particles]
                tp[s] = [p.type == PT.PROTON for p in nuc.particles]
            out[key + "_xy"] = xy
            out[key + "_isp"] = tp
    os.makedirs(os.path.join(ROOT, "pyqmd_b200", "data"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "pyqmd_b200", "data", "layout_templates.npz"), **out)
    print(f"layout templates: {len(out) // 2} isotopes x {n_seeds} seeds in {time.time() - t0:.1f}s")


# ------------------------------------------------------------------------------------------
def gen_decay_tables(zmax=110, nmax=170):
    dc = R.decay_chains
    real_random = dc.random
    rows = []
    r_values = [0.0, 0.5, 0.99, 0.995, 0.9998, 0.99985, 1.0]
    for z in range(0, zmax + 1):
        for n in range(0, nmax + 1):
            # half-life: value for u = 0, 0.5, 1 and the number of draws consumed
            hl = []
            used = 0
            for u in (0.0, 0.5, 1.0):
                fd = ref_loader.DrawFeeder([u])
                dc.random = fd
                hl.append(hx(float(dc.get_half_life(z, n))))
                used = fd.used
            # decay product for a set of branch draws
            prods = []
            pused = 0
            for r in r_values:
                fd = ref_loader.DrawFeeder([r])
                dc.random = fd
                nz, nn, mode, creator = dc.get_decay_product(z, n)
                prods.append([nz, nn, -1 if mode is None else mode.value])
                pused = fd.used
            # compress: all r give the same product for single-option nuclides
            if all(p == prods[0] for p in prods):
                prods = [prods[0]]
            rows.append([z, n, hl if used else hl[:1], used, prods, pused])
    dc.random = real_random
    with gzip.open(os.path.join(GOLD, "decay_tables.json.gz"), "wt") as f:
        json.dump(dict(source="get_half_life decay_chains.py:247-328 (u=0,0.5,1); "
                              "get_decay_product :203-245",
                       zmax=zmax, nmax=nmax, r_values=hexlist(r_values), rows=rows), f,
                  separators=(",", ":"))
    print("decay tables:", len(rows), "nuclides")


# ------------------------------------------------------------------------------------------
def ref_decay_event(nuc, feeder):
    """Physics slice of NuclearSimulation.handle_decay (nuclear_sim.py:213,215,288-294,353)
    executed with the reference's own functions; cosmetic/logging parts omitted."""
    dc = R.decay_chains
    dc.random = feeder
    p, n, decay_type, products = dc.get_decay_product(nuc.protons, nuc.neutrons)   # :213
    emitted = []
    if decay_type:                                                                 # :215
        nuc.protons = p                                                            # :288
        nuc.neutrons = n                                                           # :289
        nuc.adjust_particles(decay_type)                                           # :290
        nuc.update_center_of_mass()                                                # :291
        emitted = products(nuc.x, nuc.y)                                           # :294
        nuc.stability = dc.get_half_life(nuc.protons, nuc.neutrons)                # :353
    return decay_type, emitted


def nucleus_record(nuc):
    d = state_of(nuc.particles)
    d.update(z=nuc.protons, n=nuc.neutrons, T=hx(float(nuc.stability)), cx=hx(float(nuc.x)),
             cy=hx(float(nuc.y)))
    return d


def gen_decay_events():
    dc, pm = R.decay_chains, R.particles
    real_dc_random, real_pm_random = dc.random, pm.random
    out = {"source": "particles.py:126-208, decay_chains.py:203-421, nuclear_sim.py:161-173,"
                     "213,288-294,353"}

    # (1) should_decay probabilities: u just below / above p decides exactly
    probs = []
    Ts = [180825048000.0, 1.409993568e+17, 0.806, 164.3e-6, 3.1 * 60, 12.32 * 31557600.0,
          1e-9, 1e30, float("inf")]
    for T in Ts:
        for dt in (T * 1e-3 if math.isfinite(T) else 1.0, T * 0.01 if math.isfinite(T) else 2.0,
                   T * 0.010000001 if math.isfinite(T) else 3.0, T * 0.1 if math.isfinite(T) else 4.0,
                   T if math.isfinite(T) else 5.0, T * 50 if math.isfinite(T) else 6.0, 1 / 240,
                   4.1666e-3, 1e6):
            nuc = dc.Nucleus(6, 8, 0, 0)
            nuc.stability = T
            # bisection on the decision to recover p exactly: decision = (u < p)
            fd = ref_loader.DrawFeeder([])
            dc.random = fd
            decisions = []
            us = [0.0, 1e-300, 0.25, 0.5, 0.75, 1.0 - 2 ** -53]
            for u in us:
                fd.draws = [u]
                fd.used = 0
                decisions.append([hx(u), int(nuc.should_decay(dt)), fd.used])
            # exact p: the smallest double u for which the reference answers "no decay"
            consumed = decisions[0][2]
            p_exact = None
            if consumed:
                def dec(u):
                    fd.draws = [u]
                    fd.used = 0
                    return nuc.should_decay(dt)
                top = 1.0 - 2 ** -53
                if not dec(0.0):
                    p_exact = 0.0
                elif dec(top):
                    p_exact = 1.0
                else:
                    lo, hi = 0.0, top          # dec(lo) True, dec(hi) False
                    while True:
                        mid = (lo + hi) / 2
                        if mid == lo or mid == hi:
                            break
                        if dec(mid):
                            lo = mid
                        else:
                            hi = mid
                    p_exact = hi
            probs.append(dict(T=hx(T), dt=hx(dt), consumed=consumed,
                              p=None if p_exact is None else hx(p_exact), decisions=decisions))
    out["should_decay"] = probs

    # (2) seeded decision strings on the real MT19937 stream
    dc.random = real_dc_random
    seeded = []
    for seed, (z, n), frac, count in ((12345, (6, 8), 0.1, 64), (7, (92, 146), 0.5, 64),
                                      (99, (2, 3), 0.02, 64), (5, (6, 8), 1e-3, 256)):
        nuc = dc.Nucleus(z, n, 0, 0)
        dt = frac * nuc.stability
        random.seed(seed)
        us = [random.random() for _ in range(count)]
        random.seed(seed)
        bits = "".join("1" if nuc.should_decay(dt) else "0" for _ in range(count))
        seeded.append(dict(seed=seed, z=z, n=n, dt=hx(dt), T=hx(nuc.stability), bits=bits,
                           uniforms=hexlist(us)))
    out["seeded"] = seeded

    # (3) adjust_particles on random type lists
    adj = []
    rng = random.Random(77)
    for case in range(40):
        a = rng.choice([1, 2, 3, 4, 5, 8, 14, 30])
        types = [rng.random() < rng.choice([0.0, 0.3, 0.5, 1.0]) for _ in range(a)]
        for mode in DT:
            nuc = object.__new__(pm.Nucleus)
            nuc.particles = [pm.Particle(float(i), float(-i), PT.PROTON if t else PT.NEUTRON,
                                         1.0 + i, 2.0 - i) for i, t in enumerate(types)]
            nuc.protons, nuc.neutrons, nuc.x, nuc.y = sum(types), a - sum(types), 0.0, 0.0
            nuc.adjust_particles(mode)
            adj.append(dict(mode=mode.value, types=[int(t) for t in types],
                            out_types=[int(p.type == PT.PROTON) for p in nuc.particles],
                            out_index=[int(p.x) for p in nuc.particles],
                            out_vx=hexlist(p.vx for p in nuc.particles)))
    out["adjust"] = adj

    # (4) chain walks: forced decays (like the SPACE key, nuclear_sim.py:433-434) down the chain
    walks = []
    for (z, n), n_events, seed in (((92, 146), 22, 1), ((6, 8), 2, 2), ((1, 2), 2, 3),
                                   ((43, 56), 6, 4), ((84, 134), 6, 5), ((83, 131), 6, 6),
                                   ((47, 61), 8, 7), ((26, 33), 6, 8), ((2, 3), 5, 9),
                                   ((8, 9), 2, 10), ((79, 119), 6, 11), ((82, 127), 6, 12)):
        random.seed(seed)
        nuc = pm.Nucleus(z, n, 0.0, 0.0)
        for p in nuc.particles:                      # FP32-representable start
            p.x, p.y = float(np.float32(p.x)), float(np.float32(p.y))
        dc.random = real_dc_random
        nuc.stability = dc.get_half_life(z, n) if (z, n) in dc.HALF_LIVES else 1.0
        rng = random.Random(1000 + seed)
        events = []
        start = nucleus_record(nuc)
        for e in range(n_events):
            draws = [rng.random(), rng.random(), rng.random()]
            if (z, n) in ((84, 134), (83, 131), (43, 56)) and e == 0:
                draws[0] = 0.99999                   # exercise the rare branch once
            fd = ref_loader.DrawFeeder(draws)
            # the reference consumes draws in stream order: branch?, angle?, half-life?
            n_opts = len(dc.DECAY_CHAINS.get((nuc.protons, nuc.neutrons), [0]))
            mode, emitted = ref_decay_event(nuc, fd)
            events.append(dict(draws=hexlist(draws), used=fd.used, n_opts=n_opts,
                               mode=-1 if not mode else mode.value,
                               emitted=[[p.type.value, hx(p.x), hx(p.y), hx(p.vx), hx(p.vy)]
                                        for p in emitted],
                               after=nucleus_record(nuc)))
        walks.append(dict(z=z, n=n, start=start, events=events))
    out["walks"] = walks

    # (5) sub-step loops: decay test -> event -> force step (nuclear_sim.py:165-173)
    loops = []
    nf = R.forces()
    for (z, n), n_steps, frac, seed in (((6, 8), 12, 0.3, 21), ((2, 3), 10, 0.5, 22),
                                        ((92, 146), 6, 0.8, 23), ((84, 134), 8, 0.7, 24)):
        pm.random, dc.random = real_pm_random, real_dc_random
        random.seed(seed)
        nuc = pm.Nucleus(z, n, 0.0, 0.0)
        for p in nuc.particles:
            p.x, p.y = float(np.float32(p.x)), float(np.float32(p.y))
        nuc.stability = dc.get_half_life(z, n)
        dt_decay = frac * nuc.stability
        rng = random.Random(2000 + seed)
        start = nucleus_record(nuc)
        steps = []
        for s in range(n_steps):
            draws = [rng.random() for _ in range(4)]
            fd0 = ref_loader.DrawFeeder([draws[0]])
            pm.random = fd0
            fd = ref_loader.DrawFeeder(draws[1:])
            mode, emitted = None, []
            decayed = nuc.should_decay(dt_decay)                          # nuclear_sim.py:166
            if decayed:
                mode, emitted = ref_decay_event(nuc, fd)                  # :167
            if len(nuc.particles) > 0:                                    # :169
                nf.update_particles_cpu(nuc.particles, 1 / 240)           # :173
            steps.append(dict(draws=hexlist(draws), used0=fd0.used, used=fd.used,
                              decayed=int(decayed), mode=-1 if not mode else mode.value,
                              emitted=[[p.type.value, hx(p.x), hx(p.y), hx(p.vx), hx(p.vy)]
                                       for p in emitted],
                              after=nucleus_record(nuc)))
        loops.append(dict(z=z, n=n, dt_decay=hx(dt_decay), dt_phys=hx(1 / 240), start=start,
                          steps=steps))
    out["loops"] = loops
    dc.random, pm.random = real_dc_random, real_pm_random

    with gzip.open(os.path.join(GOLD, "decay_events.json.gz"), "wt") as f:
        json.dump(out, f, separators=(",", ":"))
    print("decay events:", len(probs), "probabilities,", len(seeded), "seeded strings,",
          len(adj), "adjust cases,", len(walks), "walks,", len(loops), "loops")


def gen_resolve_overlaps():
    """NuclearSimulation.resolve_overlaps (nuclear_sim.py:355-379) on reference nuclei: fresh
    layouts, layouts after force steps, and a frame loop (4 sub-steps + projection per frame,
    nuclear_sim.py:161-176)."""
    import importlib
    ns = importlib.import_module("nuclear_sim")
    pm = R.particles
    sim = object.__new__(ns.NuclearSimulation)
    nf = R.forces()
    cases = []

    def project(ps, draws=()):
        nuc = object.__new__(pm.Nucleus)
        nuc.particles = ps
        sim.nucleus = nuc
        fd = ref_loader.DrawFeeder(list(draws))
        ns.random = fd
        sim.resolve_overlaps()
        ns.random = random
        return fd.used

    for (z, n), seed, pre_steps in (((6, 8), 1, 0), ((2, 2), 2, 0), ((26, 30), 3, 0), ((26, 30), 3, 6),
                                    ((82, 126), 4, 0), ((82, 126), 4, 10), ((92, 146), 5, 3),
                                    ((1, 0), 6, 0)):
        random.seed(seed)
        nuc = pm.Nucleus(z, n, 0.0, 0.0)
        ps = nuc.particles
        for p in ps:
            p.x, p.y = float(np.float32(p.x)), float(np.float32(p.y))
        for _ in range(pre_steps):
            nf.update_particles_cpu(ps, 1 / 240)
        for p in ps:                              # FP32-representable input for the device side
            p.x, p.y = float(np.float32(p.x)), float(np.float32(p.y))
        inp = state_of(ps)
        used = project(ps)
        cases.append(dict(name=f"z{z}n{n}_pre{pre_steps}", input=inp, output=state_of(ps), used=used,
                          draws=[]))
    # degenerate pair (dist < 0.001) consumes one draw
    ps = R.make_particles([0.0, 0.0002, 3.0], [0.0, 0.0003, 0.5], [0] * 3, [0] * 3, [1, 0, 1])
    inp = state_of(ps)
    used = project(ps, [0.3125])
    cases.append(dict(name="degenerate", input=inp, output=state_of(ps), used=used, draws=[hx(0.3125)]))
    # frame loop: 4 force sub-steps then one projection, 12 frames (no decay)
    random.seed(9)
    nuc = pm.Nucleus(26, 30, 0.0, 0.0)
    ps = nuc.particles
    for p in ps:
        p.x, p.y = float(np.float32(p.x)), float(np.float32(p.y))
    frames = [state_of(ps)]
    for f in range(12):
        for _ in range(4):
            nf.update_particles_cpu(ps, 1 / 240)
        project(ps)
        frames.append(state_of(ps))
    with gzip.open(os.path.join(GOLD, "resolve_overlaps.json.gz"), "wt") as f:
        json.dump(dict(source="NuclearSimulation.resolve_overlaps nuclear_sim.py:355-379; frame loop "
                              ":161-176", cases=cases, frames=frames, substeps_per_frame=4,
                       dt=hx(1 / 240)), f, separators=(",", ":"))
    print("resolve_overlaps:", len(cases), "cases,", len(frames) - 1, "frames")


def gen_sim_driver():
    """Frame-level host logic of NuclearSimulation: the sub-step plan of update_simulation
    (nuclear_sim.py:123-153), the cosmetic speed / lifetime rewrite of emitted particles in
    handle_decay (:295-342) and the free-particle animation update_particle (:178-210)."""
    import importlib
    from collections import deque
    ns = importlib.import_module("nuclear_sim")
    pm, dc = R.particles, R.decay_chains
    out = {"source": "nuclear_sim.py:118-176,178-210,295-347"}

    def fresh():
        sim = object.__new__(ns.NuclearSimulation)
        sim.fps_history = deque(maxlen=30)
        sim.time_scale, sim.time_passed = 1.0, 0
        sim.camera_pos, sim.camera_target = [400, 400], [400, 400]
        sim.zoom_level = sim.target_zoom = 15.0
        sim.zoom_speed = 0.1
        sim.auto_adjust_substeps, sim.physics_dt_factor = False, 0.8
        sim.physics_dt, sim.accuracy, sim.max_substeps, sim.substeps_used = 1 / 240, 1, 20, 0
        sim.nucleus, sim.particles, sim.gpu_available = None, [], False
        sim.decay_times = deque(maxlen=100)
        return sim

    plans = []
    ns.random = ref_loader.DrawFeeder([0.5] * 100000)
    for auto in (False, True):
        for ts in (1e-3, 0.5, 1.0, 2.0, 10.0, 60.0, 3600.0, 31557600.0, 3.15576e16):
            for dt in (1 / 240, 1 / 144, 1 / 60, 1 / 30, 0.1, 0.5):
                sim = fresh()
                sim.auto_adjust_substeps, sim.time_scale = auto, ts
                sim.update_simulation(dt)
                plans.append(dict(auto=auto, time_scale=hx(ts), dt=hx(dt), num_steps=sim.substeps_used,
                                  physics_dt=hx(sim.physics_dt), time_passed=hx(float(sim.time_passed))))
    out["plans"] = plans

    cosmetics = []
    for ts in (0.5, 1.0, 60.0, 86400.0, 31557600000.0):
        for sub in (1, 4, 12, 16, 20):
            for pdt in (1 / 240, 1 / 60, 1 / 1000):
                for (z, n) in ((92, 146), (6, 8), (84, 134), (43, 56)):
                    sim = fresh()
                    sim.time_scale, sim.substeps_used, sim.physics_dt = ts, sub, pdt
                    random.seed(z)
                    nuc = pm.Nucleus(z, n, 0.0, 0.0)
                    nuc.stability = 1.0
                    nuc.decay_chain, nuc.last_decay_time = [], 0.0
                    sim.nucleus, sim.time_passed = nuc, 10.0
                    dc.random = ref_loader.DrawFeeder([0.7, 0.3, 0.4])
                    ns.random = ref_loader.DrawFeeder([0.5] * 16)
                    sim.handle_decay()
                    for p in sim.particles:
                        cosmetics.append(dict(time_scale=hx(ts), substeps=sub, physics_dt=hx(pdt),
                                              ptype=p.type.value, vx=hx(p.vx), vy=hx(p.vy),
                                              lifetime=hx(float(p.lifetime))))
    dc.random = random
    out["cosmetics"] = cosmetics

    anim = []
    for ts in (1.0, 50.0, 1e4):
        for sub in (1, 4, 20):
            for ptype in (pm.ParticleType.ALPHA, pm.ParticleType.ELECTRON, pm.ParticleType.GAMMA,
                          pm.ParticleType.POSITRON, pm.ParticleType.NEUTRON, pm.ParticleType.PROTON):
                sim = fresh()
                sim.time_scale, sim.substeps_used = ts, sub
                p = pm.Particle(1.0, -2.0, ptype, 30.0, -40.0)
                p.lifetime = 0.05 if ptype != pm.ParticleType.NEUTRON else p.lifetime
                alive = []
                for k in range(6):
                    alive.append(bool(sim.update_particle(p, 1 / 240, 0.004 * (k + 1))))
                anim.append(dict(time_scale=hx(ts), substeps=sub, ptype=ptype.value, x=hx(p.x),
                                 y=hx(p.y), age=hx(float(p.age)), alive=alive))
    out["animation"] = anim
    ns.random = random
    with gzip.open(os.path.join(GOLD, "sim_driver.json.gz"), "wt") as f:
        json.dump(out, f, separators=(",", ":"))
    print("sim driver:", len(plans), "plans,", len(cosmetics), "cosmetic cases,", len(anim), "animation cases")


def check_handle_decay_slice():
    """Sanity: the physics slice used above equals the real NuclearSimulation.handle_decay
    (nuclear_sim.py:212-353) in Z, N, particle list, centre and stability."""
    import importlib
    ns = importlib.import_module("nuclear_sim")
    dc, pm = R.decay_chains, R.particles
    real = dc.random
    sim = object.__new__(ns.NuclearSimulation)
    sim.particles, sim.time_passed, sim.substeps_used = [], 10.0, 4
    sim.time_scale, sim.physics_dt = 1.0, 1 / 240
    from collections import deque
    sim.decay_times = deque(maxlen=100)
    ok = True
    for seed in range(3):
        random.seed(seed)
        a = pm.Nucleus(92, 146, 400, 400)
        random.seed(seed)
        b = pm.Nucleus(92, 146, 400, 400)
        a.stability = b.stability = dc.get_half_life(92, 146)
        a.decay_chain, a.last_decay_time = [], 0.0
        sim.nucleus = a
        for e in range(5):
            draws = [0.3 + 0.1 * e, 0.6, 0.2]
            dc.random = ref_loader.DrawFeeder(draws)
            ns.random = ref_loader.DrawFeeder([0.5] * 8)     # cosmetic draws :251,:345
            sim.handle_decay()
            ref_decay_event(b, ref_loader.DrawFeeder(draws))
            same = (a.protons, a.neutrons, a.stability, a.x, a.y) == \
                   (b.protons, b.neutrons, b.stability, b.x, b.y)
            same = same and [(p.x, p.y, p.vx, p.vy, p.type) for p in a.particles] == \
                [(p.x, p.y, p.vx, p.vy, p.type) for p in b.particles]
            ok = ok and same
    dc.random = real
    ns.random = random
    print("handle_decay physics slice == real handle_decay:", ok)
    assert ok


if __name__ == "__main__":
    what = sys.argv[1:] or ["data", "kats", "tables", "events", "check", "overlaps", "driver", "layouts", "traj"]
    if "data" in what:
        gen_nuclide_data()
    if "kats" in what:
        gen_force_kats()
    if "tables" in what:
        gen_decay_tables()
    if "events" in what:
        gen_decay_events()
    if "check" in what:
        check_handle_decay_slice()
    if "overlaps" in what:
        gen_resolve_overlaps()
    if "driver" in what:
        gen_sim_driver()
    if "layouts" in what:
        gen_layouts()
    if "traj" in what:
        gen_u238_traj()
