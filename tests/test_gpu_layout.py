"""Device-side initial layout (pyqmd_ensemble_init_layout) against the reference's placement
algorithm Nucleus.initialize_particles (particles.py:62-124).  The host mirror
pyqmd_b200.types.Nucleus reproduces the reference bit for bit (tests/test_host.py); here both sides
are fed the same uniform draws."""
import numpy as np
import pytest
import torch

from pyqmd_b200 import types as T
from pyqmd_b200.state import NucleusEnsemble, README_ISOTOPES, layout_templates

pytestmark = pytest.mark.gpu


class Feeder:
    def __init__(self, draws):
        self.draws, self.used = draws, 0

    def random(self):
        self.used += 1
        return float(self.draws[self.used - 1])

    def uniform(self, a, b):
        return a + (b - a) * self.random()


def host_layout(z, n, draws, monkeypatch):
    monkeypatch.setattr(T, "random", Feeder(draws.reshape(-1)))
    nuc = T.Nucleus(z, n, 0.0, 0.0)
    xy = np.array([[p.x, p.y] for p in nuc.particles])
    isp = np.array([p.type == T.ParticleType.PROTON for p in nuc.particles], np.uint8)
    return xy, isp


@pytest.mark.parametrize("zn", [(1, 0), (1, 2), (2, 2), (6, 8), (26, 30), (47, 60), (92, 146), (3, 1)])
def test_injected_draws_match_reference_placement(zn, monkeypatch):
    z, n = zn
    a = z + n
    rng = np.random.default_rng(100 * z + n)
    n_nuc = 3
    draws = rng.random((n_nuc, a, 21))
    ens = NucleusEnsemble.from_device_layout((zn,), n_nuc, decay=False, layout_uniforms={a: draws})
    pos = ens.pos.cpu().numpy().reshape(n_nuc, a, 2)
    isp = ens.is_proton.cpu().numpy().reshape(n_nuc, a)
    assert (ens.count.cpu().numpy() == a).all()
    assert float(ens.vel.abs().max()) == 0.0
    for k in range(n_nuc):
        xy, tp = host_layout(z, n, draws[k], monkeypatch)
        assert np.array_equal(isp[k], tp), "placement order (p/n pairs per shell, then the surplus)"
        want = xy.astype(np.float32)
        # float64 cos/sin of CUDA and glibc may differ in the last bit: <= 1 FP32 ulp after rounding
        tol = np.maximum(np.abs(want), 1e-3) * 2.0 ** -22
        assert (np.abs(pos[k] - want) <= tol).all(), (zn, k, np.abs(pos[k] - want).max())
        assert (pos[k] == want).mean() > 0.98


def test_philox_layouts_have_the_reference_statistics():
    """Without injected draws: every nucleus its own layout; shells and same-type spacing look like
    the reference-generated templates."""
    n_nuc = 512
    ens = NucleusEnsemble.from_device_layout(((82, 126),), n_nuc, decay=False, layout_seed=7)
    pos = ens.pos.cpu().numpy().reshape(n_nuc, 208, 2)
    isp = ens.is_proton.cpu().numpy().reshape(n_nuc, 208)
    tm = layout_templates()
    txy, tis = tm["z82_n126_xy"], tm["z82_n126_isp"]
    assert np.array_equal(isp[0], tis[0]) and (isp == isp[0]).all()        # order depends on (Z, N) only
    assert len({pos[k].tobytes() for k in range(n_nuc)}) == n_nuc          # all different
    r_dev, r_ref = np.hypot(pos[..., 0], pos[..., 1]), np.hypot(txy[..., 0], txy[..., 1])
    # radius of nucleon k: shell radius * U(0.8, 1): same support, same mean per slot
    assert np.abs(r_dev.mean(0) - r_ref.mean(0)).max() < 0.12 * r_ref.mean(0).max()
    assert r_dev.max() <= r_ref.max() * 1.02 and r_dev.min() >= r_ref.min() * 0.9

    def nn_same(p, t):
        out = []
        for k in range(p.shape[0]):
            for kind in (0, 1):
                q = p[k][t[k] == kind]
                d = np.hypot(q[:, None, 0] - q[None, :, 0], q[:, None, 1] - q[None, :, 1])
                np.fill_diagonal(d, 1e9)
                out.append(d.min(1).mean())
        return float(np.mean(out))
    assert abs(nn_same(pos[:64], isp[:64]) - nn_same(txy, tis)) < 0.05 * nn_same(txy, tis)
    # a different seed gives different layouts, the same seed the same ones
    again = NucleusEnsemble.from_device_layout(((82, 126),), 8, decay=False, layout_seed=7)
    other = NucleusEnsemble.from_device_layout(((82, 126),), 8, decay=False, layout_seed=8)
    assert torch.equal(again.pos, ens.pos[: 8 * 208])
    assert not torch.equal(other.pos, ens.pos[: 8 * 208])


def test_mixed_ensemble_device_layout_steps():
    ens = NucleusEnsemble.from_device_layout(README_ISOTOPES, 9 * 20, decay=True, dt_decay=1.0,
                                             layout_seed=3, seed=1)
    cnt = ens.count.cpu().numpy()
    assert sorted(set(cnt.tolist())) == sorted({z + n for z, n in README_ISOTOPES})
    ens.frame(4)
    assert torch.isfinite(ens.pos).all()


def test_random_nuclides_layout_matches_reference_placement(monkeypatch):
    """Randomised: 16 random (Z, N) with up to 90 nucleons (surplus protons, surplus neutrons, Z = 0
    or N = 0), injected draws, against the host mirror of the reference placement."""
    rng = np.random.default_rng(4)
    for _ in range(16):
        a = int(rng.integers(1, 91))
        z = int(rng.integers(0, a + 1))
        n = a - z
        draws = rng.random((1, a, 21))
        ens = NucleusEnsemble.from_device_layout(((z, n),), 1, decay=False, layout_uniforms={a: draws})
        xy, tp = host_layout(z, n, draws[0], monkeypatch)
        assert np.array_equal(ens.is_proton.cpu().numpy(), tp), (z, n)
        want = xy.astype(np.float32)
        tol = np.maximum(np.abs(want), 1e-3) * 2.0 ** -22
        assert (np.abs(ens.pos.cpu().numpy() - want) <= tol).all(), (z, n)
