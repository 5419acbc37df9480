"""Host-side frame logic of pyqmd_b200.sim against golden vectors generated from the reference's
NuclearSimulation (tests/golden/gen_golden.py::gen_sim_driver): the sub-step plan of
update_simulation (nuclear_sim.py:123-153), the speed / lifetime rewrite of emitted particles
(:295-342) and the free-particle animation update_particle (:178-210).  CPU only, bit-exact."""
import numpy as np
import pytest

from conftest import load_json
from pyqmd_b200 import sim


@pytest.fixture(scope="module")
def golden():
    return load_json("sim_driver.json.gz")


def fh(s):
    return float.fromhex(s)


def test_substep_plan_bit_exact(golden):
    for row in golden["plans"]:
        ts, dt = fh(row["time_scale"]), fh(row["dt"])
        num_steps, eff, step_time, physics_dt = sim.substep_plan(
            dt, ts, auto_adjust_substeps=row["auto"])
        assert num_steps == row["num_steps"], row
        assert physics_dt == fh(row["physics_dt"]), row
        assert eff == physics_dt * 1.0
        assert step_time == dt * ts / num_steps
        assert dt * ts == fh(row["time_passed"])


def test_default_realtime_frame_is_four_substeps():
    # 60 fps real time -> 4 sub-steps of 1/240
    assert sim.substep_plan(1 / 60, 1.0)[:2] == (4, 1 / 240)
    # the cap: max_substeps = 20 (nuclear_sim.py:63)
    assert sim.substep_plan(1 / 60, 3.15576e16)[0] == 20


def test_cosmetic_speed_and_lifetime_bit_exact(golden):
    from pyqmd_b200.types import ParticleType
    speeds = {ParticleType.ALPHA.value: 30.0, ParticleType.GAMMA.value: 60.0,
              ParticleType.ELECTRON.value: 50.0, ParticleType.POSITRON.value: 50.0}
    seen = set()
    for row in golden["cosmetics"]:
        # lifetime depends on (time scale, sub-steps, physics dt) only; re-normalising an already
        # renormalised velocity must keep the base speed
        vx, vy = fh(row["vx"]), fh(row["vy"])
        gx, gy, life = sim.cosmetic_speed_lifetime(row["ptype"], vx, vy, fh(row["time_scale"]),
                                                   row["substeps"], fh(row["physics_dt"]))
        assert life == fh(row["lifetime"]), row
        # renormalising an already renormalised velocity keeps speed = base speed
        assert abs(np.hypot(gx, gy) - speeds.get(row["ptype"], 40.0)) < 1e-9
        assert abs(np.hypot(vx, vy) - speeds.get(row["ptype"], 40.0)) < 1e-9
        seen.add(row["ptype"])
    assert {ParticleType.ALPHA.value, ParticleType.ELECTRON.value} <= seen


class _Feeder:
    """Stands in for the ``random`` module: serves supplied draws in order."""

    def __init__(self, draws):
        self.draws, self.used = list(draws), 0

    def random(self):
        self.used += 1
        return self.draws[self.used - 1]

    def uniform(self, a, b):
        return a + (b - a) * self.random()


def test_emission_plus_rewrite_matches_reference_bit_exact(golden, monkeypatch):
    """get_decay_product -> creator(x, y) (decay_chains.py:203-245, 331-371) followed by the
    rewrite reproduces the (vx, vy) the reference's handle_decay left on the emitted particle,
    bit for bit, when fed the same draws (golden rows cycle over four isotopes; draw order:
    [branch], angle, [daughter half-life])."""
    from pyqmd_b200 import nuclides
    isotopes = ((92, 146), (6, 8), (84, 134), (43, 56))
    assert len(golden["cosmetics"]) % len(isotopes) == 0
    for i, row in enumerate(golden["cosmetics"]):
        z, n = isotopes[i % len(isotopes)]
        monkeypatch.setattr(nuclides, "random", _Feeder([0.7, 0.3, 0.4]))
        _, _, mode, creator = nuclides.get_decay_product(z, n)
        ps = creator(0.0, 0.0)
        assert len(ps) == 1 and ps[0].type.value == row["ptype"], (z, n, mode)
        vx, vy, life = sim.cosmetic_speed_lifetime(row["ptype"], ps[0].vx, ps[0].vy,
                                                   fh(row["time_scale"]), row["substeps"],
                                                   fh(row["physics_dt"]))
        assert vx == fh(row["vx"]) and vy == fh(row["vy"]), (i, z, n)
        assert life == fh(row["lifetime"])


def test_animation_bit_exact(golden):
    from pyqmd_b200.types import ParticleType
    for row in golden["animation"]:
        ts, sub, ptype = fh(row["time_scale"]), row["substeps"], row["ptype"]
        x, y, age = np.array([1.0]), np.array([-2.0]), np.array([0.0])
        life = np.array([0.05 if ptype != ParticleType.NEUTRON.value else np.inf])
        alive = []
        for k in range(6):
            x, y, age, a = sim.animate(np.array([ptype]), x, y, np.array([30.0]), np.array([-40.0]),
                                       age, life, 1 / 240, 0.004 * (k + 1), ts, sub)
            alive.append(bool(a[0]))
        assert alive == row["alive"], row
        assert x[0] == fh(row["x"]) and y[0] == fh(row["y"]), row
        assert age[0] == fh(row["age"]), row


def test_time_scale_presets_match_reference():
    assert sim.TIME_SCALE_PRESETS["billion"] == 31557600000000000.0
    assert sim.TIME_SCALE_PRESETS["real"] == 1.0
    assert len(sim.TIME_SCALE_PRESETS) == 8


def test_vectorised_cosmetics_equal_the_scalar_function_bit_for_bit(golden):
    rng = np.random.default_rng(3)
    n = 4000
    ptype = rng.integers(0, 6, n)
    ang = rng.random(n) * 2 * np.pi
    speed = np.where(rng.random(n) < 0.02, 0.0, rng.random(n) * 300)
    vx, vy = speed * np.cos(ang), speed * np.sin(ang)
    for ts, sub, pdt in ((0.5, 3, 1 / 240), (1.0, 4, 1 / 240), (60.0, 16, 1 / 60), (3.15576e16, 20, 1 / 1000)):
        ax, ay, al = sim.cosmetic_speed_lifetime_array(ptype, vx, vy, ts, sub, pdt)
        for k in range(n):
            bx, by, bl = sim.cosmetic_speed_lifetime(int(ptype[k]), float(vx[k]), float(vy[k]), ts, sub, pdt)
            assert ax[k] == bx and ay[k] == by and al[k] == bl, (k, ts)
    rows = golden["cosmetics"]
    gx, gy, gl = sim.cosmetic_speed_lifetime_array(
        np.array([r["ptype"] for r in rows[:20]]), np.array([fh(r["vx"]) for r in rows[:20]]),
        np.array([fh(r["vy"]) for r in rows[:20]]), fh(rows[0]["time_scale"]), rows[0]["substeps"],
        fh(rows[0]["physics_dt"]))
    assert gl[0] == fh(rows[0]["lifetime"])
