"""Soak run of the headless frame driver: many frames with heavy decay activity; checks that the
state stays finite and the bookkeeping consistent (run through gpurun)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyqmd_b200 import nuclides
from pyqmd_b200.sim import HeadlessSimulation, TIME_SCALE_PRESETS
from pyqmd_b200.state import CODE_ISOTOPES

out = {}
for name, kw, frames in (("u238_x2048", dict(isotope=(92, 146), n_nuclei=2048, time_scale=TIME_SCALE_PRESETS["billion"] * 300), 300),
                         ("code_isotopes_x4500", dict(isotopes=CODE_ISOTOPES, n_nuclei=4500, time_scale=TIME_SCALE_PRESETS["year"]), 300)):
    sim = HeadlessSimulation(seed=3, **kw)
    t0 = time.perf_counter()
    last = 0
    for f in range(frames):
        sim.update_simulation(1 / 60)
        tot = sum(sim.decay_counts.values())
        assert tot >= last
        last = tot
    torch.cuda.synchronize()
    ens = sim.ensemble
    cnt = ens.count.cpu().numpy()
    zn = ens.zn.cpu().numpy()
    off = ens.offsets.cpu().numpy()
    assert torch.isfinite(ens.pos).all() and torch.isfinite(ens.vel).all()
    # live nucleon lists follow (Z, N) on alpha / beta chains (sample)
    isp = ens.is_proton.cpu().numpy()
    bad = 0
    for k in range(0, ens.n_nuclei, 97):
        z, n = nuclides.zn_unpack(int(zn[k]))
        p = int(isp[off[k]:off[k] + cnt[k]].sum())
        bad += (p != z) or (cnt[k] - p != n)
    pos = ens.pos.cpu().numpy()
    ext = [float(np.hypot(*(pos[off[k]:off[k] + cnt[k]] - pos[off[k]:off[k] + cnt[k]].mean(0)).T).max())
           for k in range(0, ens.n_nuclei, 211) if cnt[k] > 1]
    out[name] = {"frames": frames, "seconds": time.perf_counter() - t0, "decays": sim.decay_counts,
                 "events_seen": sim._events_seen, "events_dropped": sim.events_dropped,
                 "free_particles": int(len(sim.free["x"])), "min_count": int(cnt.min()),
                 "list_vs_zn_mismatches_in_sample": int(bad), "extent_min_max": [min(ext), max(ext)],
                 "substeps_used": sim.substeps_used}
print(json.dumps(out))
