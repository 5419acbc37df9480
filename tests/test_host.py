"""Host-side logic of the product (no GPU): the reference-shaped decay interface, the device
table builder, the domain types, the C-ABI library's exports."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

from conftest import ROOT, unhex
from pyqmd_b200 import _lib, nuclides
from pyqmd_b200.types import DecayType, Nucleus, Particle, ParticleType


class Feeder:
    """Replaces the module-level ``random`` of pyqmd_b200.nuclides with explicit draws."""

    def __init__(self, draws):
        self.draws, self.used = list(draws), 0

    def random(self):
        self.used += 1
        return self.draws[self.used - 1]

    def uniform(self, a, b):
        return a + (b - a) * self.random()


@pytest.fixture
def feed(monkeypatch):
    def install(draws):
        f = Feeder(draws)
        monkeypatch.setattr(nuclides, "random", f)
        return f
    return install


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "pyqmd_b200.h")).read()
    declared = set(re.findall(r"\b(pyqmd_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.pyqmd_abi_version() == 1


def test_struct_layouts_match_header():
    sizes = (C.c_int64 * 4)()
    assert _lib.lib().pyqmd_struct_sizes(sizes) == 0
    assert list(sizes) == [_lib.NUCLIDE_DTYPE.itemsize, _lib.EVENT_DTYPE.itemsize,
                           C.sizeof(_lib.EnsembleDesc), C.sizeof(_lib.PopulationDesc)]


def test_argument_validation_without_gpu():
    """Error behaviour of the ABI: bad arguments give negative codes + a message, never a crash;
    empty inputs return 0 immediately (nuclear_forces.py:186-188, 238-239)."""
    L = _lib.lib()
    assert L.pyqmd_update_forces_and_positions(None, None, 0, 0, 0, 150, 30, 35, 0.004) == 0
    assert L.pyqmd_update_particles_f64(None, None, None, None, None, 0, 150, 30, 35, 0.004, 1) == 0
    assert L.pyqmd_update_forces_and_positions(None, None, 5, 0, 0, 150, 30, 35, 0.004) == -1
    assert b"NULL" in L.pyqmd_last_error()
    assert L.pyqmd_cloud_step(None, None, None, None, None, 0, 0, 0, 150, 30, 35, 0.004, None, None) == 0
    assert L.pyqmd_cloud_step(None, None, None, None, None, 10, 5, 3, 150, 30, 35, 0.004, None, None) == -1
    assert L.pyqmd_ensemble_step(None, 1, None) == -1
    assert L.pyqmd_cloud_workspace_bytes(1_000_000) > 3907 * 36


def test_half_life_and_decay_product_against_reference_dump(decay_tables, feed):
    rvals = [float.fromhex(h) for h in decay_tables["r_values"]]
    saved = {k: list(v) for k, v in nuclides.DECAY_CHAINS.items()}
    try:
        for z, n, hl, used, prods, pused in decay_tables["rows"]:
            if used:
                for u, want in zip((0.0, 0.5, 1.0), hl):
                    f = feed([u])
                    assert nuclides.get_half_life(z, n) == float.fromhex(want), (z, n, u)
                    assert f.used == 1
            else:
                f = feed([])
                assert nuclides.get_half_life(z, n) == float.fromhex(hl[0]) and f.used == 0
            for k, r in enumerate(rvals):
                want = prods[k] if len(prods) > 1 else prods[0]
                f = feed([r])
                nz, nn, mode, creator = nuclides.get_decay_product(z, n)
                assert [nz, nn, -1 if mode is None else mode.value] == want, (z, n, r)
                assert f.used == pused
                assert callable(creator)
    finally:
        nuclides.DECAY_CHAINS.clear()
        nuclides.DECAY_CHAINS.update(saved)


def test_creators(decay_events, feed):
    seen = set()
    for walk in decay_events["walks"]:
        for ev in walk["events"]:
            if not ev["emitted"]:
                continue
            draws = [float.fromhex(h) for h in ev["draws"]]
            u_angle = draws[1] if ev["n_opts"] > 1 else draws[0]
            ptype, x, y, vx, vy = ev["emitted"][0]
            creator = nuclides._CREATORS[DecayType(ev["mode"])]
            feed([u_angle])
            p, = creator(float.fromhex(x), float.fromhex(y))
            assert (p.type.value, p.x, p.y, p.vx, p.vy) == (
                ptype, float.fromhex(x), float.fromhex(y), float.fromhex(vx), float.fromhex(vy))
            seen.add(ev["mode"])
    assert {1, 2, 3, 4, 6} <= seen          # alpha, beta-, beta+, gamma, proton emission


def test_should_decay_matches_reference(decay_events, monkeypatch):
    import pyqmd_b200.types as types_mod
    for row in decay_events["should_decay"]:
        T, dt = float.fromhex(row["T"]), float.fromhex(row["dt"])
        nuc = Nucleus(6, 8, 0, 0, particles=[])
        nuc.stability = T
        for uh, dec, used in row["decisions"]:
            f = Feeder([float.fromhex(uh)])
            monkeypatch.setattr(types_mod, "random", f)
            assert nuc.should_decay(dt) == bool(dec) and f.used == used
    for row in decay_events["seeded"]:
        nuc = Nucleus(row["z"], row["n"], 0, 0, particles=[])
        nuc.stability = float.fromhex(row["T"])
        monkeypatch.setattr(types_mod, "random", random)
        random.seed(row["seed"])
        bits = "".join("1" if nuc.should_decay(float.fromhex(row["dt"])) else "0"
                       for _ in range(len(row["bits"])))
        assert bits == row["bits"]


def test_adjust_particles_matches_reference(decay_events):
    for row in decay_events["adjust"]:
        ps = [Particle(float(i), float(-i), ParticleType.PROTON if t else ParticleType.NEUTRON,
                       1.0 + i, 2.0 - i) for i, t in enumerate(row["types"])]
        nuc = Nucleus(0, 0, 0.0, 0.0, particles=ps)
        nuc.adjust_particles(DecayType(row["mode"]))
        assert [int(p.type == ParticleType.PROTON) for p in nuc.particles] == row["out_types"]
        assert [int(p.x) for p in nuc.particles] == row["out_index"]
        assert np.array_equal(np.array([p.vx for p in nuc.particles]), unhex(row["out_vx"]))


def test_initial_layout_matches_reference_templates():
    """Nucleus.initialize_particles reproduces the reference's layouts (same global-random
    consumption) -- compared with the reference-generated FP32 templates."""
    from pyqmd_b200.state import layout_templates
    tm = layout_templates()
    for (z, n), seed in (((6, 8), 0), ((2, 2), 3), ((1, 0), 5), ((26, 30), 1)):
        random.seed(1000 * (z * 256 + n) + seed)
        nuc = Nucleus(z, n, 0.0, 0.0)
        xy = np.array([[p.x, p.y] for p in nuc.particles], np.float32)
        assert np.array_equal(xy, tm[f"z{z}_n{n}_xy"][seed])
        isp = np.array([p.type == ParticleType.PROTON for p in nuc.particles], np.uint8)
        assert np.array_equal(isp, tm[f"z{z}_n{n}_isp"][seed])


def test_device_table_rows(decay_tables):
    dt = 12345.678
    tab = nuclides.build_device_table(dt)
    assert tab.shape == (_lib.TABLE_ZDIM * _lib.TABLE_NDIM,)
    from oracle import oracle as orc
    for z, n, hl, used, prods, pused in decay_tables["rows"][::7]:
        row = tab[z * _lib.TABLE_NDIM + n]
        if used:
            assert row["kind"] == _lib.HL_BAND
            for u, want in zip((0.0, 0.5, 1.0), hl):
                got = nuclides.half_life_from_draw(row["band_a"], row["band_b"], row["band_unit"], u)
                assert got == float.fromhex(want)
        else:
            T = float.fromhex(hl[0])
            assert row["half_life"] == T
            assert row["p_decay"] == orc.decay_probability(T, dt)
        assert row["n_opt"] == (2 if pused else 1)
        first = prods[0]
        if first[2] == -1:
            assert row["opt_mode"][0] == 0
        elif first[0] < 0 or first[1] < 0:
            pass        # unphysical daughters (Z = 0 parents) are clamped into the table
        else:
            assert nuclides.zn_unpack(row["opt_zn"][0]) == (first[0], first[1])
            assert row["opt_mode"][0] == first[2]
    po = tab[84 * _lib.TABLE_NDIM + 134]      # Po-218: 0.9998 + 0.0002 == 1.0 exactly
    assert list(po["opt_cum"]) == [0.9998, 1.0]


def test_shard_range():
    from pyqmd_b200.state import shard_range
    for n in (0, 1, 7, 8, 1000, 65536):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_product_does_not_import_oracle():
    """The shipped package must never route through oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "pyqmd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "liboracle" not in src, f


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "SO_PATH", "/nonexistent/libpyqmd_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_dropin_recognises_protons_of_a_foreign_particle_class():
    """nuclear_sim.py keeps building particles with the reference's own ParticleType when only
    nuclear_forces is swapped: the drop-in must read the enum VALUE (0 = proton, particles.py:6)."""
    import enum

    from pyqmd_b200 import forces

    class ForeignType(enum.Enum):
        PROTON = 0
        NEUTRON = 1

    class P:
        def __init__(self, t):
            self.type = t
    assert forces._is_proton(P(ForeignType.PROTON)) and not forces._is_proton(P(ForeignType.NEUTRON))
    assert forces._is_proton(P(forces.ParticleType.PROTON)) and not forces._is_proton(P(forces.ParticleType.NEUTRON))
    assert forces._is_proton(P(0)) and not forces._is_proton(P(1))


def test_integer_decay_threshold_is_the_same_predicate_as_random_lt_p():
    """The decay-only kernel decides `random() < p` (particles.py:147) in integers: CPython's random()
    is m / 2^53 for a 53-bit integer m, the table carries ceil(p * 2^53)."""
    from pyqmd_b200 import nuclides
    rng = np.random.default_rng(5)
    p = np.concatenate([10.0 ** rng.uniform(-25, 0, 4000), rng.random(2000), [0.0, 1.0, 2.0 ** -53, 2.0 ** -60,
                       float.fromhex("0x1.6b54e2b063e07p-11"), 0.5, 1.0 - 2.0 ** -53]])
    thr = nuclides.decay_thresholds(p)
    for k in range(len(p)):
        t = int(thr[k])
        cands = {0, 1, (1 << 53) - 1, max(t - 1, 0), t, min(t + 1, (1 << 53) - 1), int(rng.integers(0, 1 << 53))}
        for m in cands:
            assert ((m / 9007199254740992.0) < p[k]) == (m < t), (p[k], m, t)
    assert nuclides.decay_thresholds(np.array([-1.0, np.nan]))[0] == 0
    tab = nuclides.build_device_table(180825048000.0 * 1e-3)
    band = tab["kind"] == _lib_mod().HL_BAND
    assert (tab["p_thr"][band] == _lib_mod().THR_PER_NUCLEUS).all() and (tab["p_thr"][tab["p_decay"] < 0] == 0).all()


def _lib_mod():
    from pyqmd_b200 import _lib
    return _lib
