"""The reference's own Renderer (rendering.py, unmodified, from baseline/_ref) draws a GPU-resident
nucleus: HeadlessSimulation.render_args(k, types=<reference particles module>) feeds
Renderer.render (rendering.py:32-58) through a RECORDING stand-in for pygame (no display in the
container): every nucleon and every live emitted particle arrives as a pygame.draw.circle call at the
screen position world_to_screen gives it (rendering.py:121-127)."""
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class _Surface:
    def __init__(self, w=1200, h=800):
        self.w, self.h, self.ops = w, h, []

    def get_width(self):
        return self.w

    def get_height(self):
        return self.h

    def fill(self, color, rect=None):
        self.ops.append(("fill", color))

    def blit(self, src, pos, *a):
        self.ops.append(("blit", pos))


class _Font:
    def render(self, text, aa, color):
        s = _Surface(8 * len(str(text)), 16)
        s.text = str(text)
        return s

    def size(self, text):
        return 8 * len(str(text)), 16

    def get_height(self):
        return 16


def recording_pygame(mod):
    calls = {"circle": [], "line": [], "rect": [], "flip": 0, "text": []}
    mod.font = types.SimpleNamespace(SysFont=lambda *a, **k: _Font())
    mod.draw = types.SimpleNamespace(
        circle=lambda surf, color, pos, radius, width=0: calls["circle"].append((color, pos, radius, width)),
        line=lambda surf, color, a, b, width=1: calls["line"].append((a, b)),
        rect=lambda surf, color, rect, width=0, **k: calls["rect"].append(rect))
    mod.display = types.SimpleNamespace(flip=lambda: calls.__setitem__("flip", calls["flip"] + 1))
    mod.Rect = lambda *a: tuple(a)
    mod.Surface = lambda size, *a, **k: _Surface(*size)
    mod.SRCALPHA = 0
    return calls


def test_reference_renderer_draws_the_gpu_resident_nucleus():
    import sys

    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("baseline/_ref not installed (python baseline/install_reference.py)")
    R = ref_loader.Ref()
    import importlib
    rendering = importlib.import_module("rendering")                # the reference's, unmodified
    calls = recording_pygame(sys.modules["pygame"])
    from pyqmd_b200 import nuclides
    from pyqmd_b200.sim import HeadlessSimulation
    sim = HeadlessSimulation((6, 8), n_nuclei=64, seed=2)
    T = nuclides.get_half_life(6, 8)
    sim.time_scale = T / 5 * 60                                     # a fifth of a half-life per frame
    for _ in range(4):
        sim.update_simulation(1 / 60)
    k = int(np.nonzero(sim.ensemble.zn.cpu().numpy() == nuclides.zn_pack(7, 7))[0][0])     # a decayed nucleus
    renderer = rendering.Renderer(_Surface())
    args = sim.render_args(k, types=R.particles)
    nucleus, particles = args[0], args[1]
    assert type(nucleus) is R.particles.Nucleus and type(nucleus.particles[0]) is R.particles.Particle
    renderer.render(*args)
    assert calls["flip"] == 1
    # every nucleon is a filled circle at world_to_screen(x, y) (camera at the origin (400, 400), zoom 15)
    cam, zoom = args[2], args[3]
    want = {(int(renderer.simulation_width / 2 + (p.x - cam[0]) * zoom),
             int(renderer.simulation_height / 2 + (p.y - cam[1]) * zoom)) for p in nucleus.particles}
    filled = {c[1] for c in calls["circle"] if c[3] == 0}
    assert len(nucleus.particles) == 14 and want <= filled
    # protons get their highlight, neutrons their ring: the renderer recognised ITS enum (rendering.py:72,81)
    n_p = sum(1 for p in nucleus.particles if p.type == R.particles.ParticleType.PROTON)
    assert n_p == 7
    rings = [c for c in calls["circle"] if c[3] == 1]
    assert len(rings) == len(nucleus.particles) - n_p
    # the emitted electron is drawn too, faded by age / lifetime (rendering.py:47)
    assert len(particles) == 1 and particles[0].type == R.particles.ParticleType.ELECTRON
    # the info panel printed the nuclide (rendering.py:157-170 reads .protons / .neutrons / .stability)
    assert calls["circle"] and len(calls["line"]) >= 7
