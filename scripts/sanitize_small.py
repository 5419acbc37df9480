"""Small invocation of every kernel for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bench import make_cloud
from pyqmd_b200.state import CODE_ISOTOPES, DecayPopulation, NucleonCloud, NucleusEnsemble, README_ISOTOPES

ens = NucleusEnsemble.from_templates(README_ISOTOPES, 18, decay=True, dt_decay=1e17, seed=3)
ens.step(3)
ens.frame(2)
ens.census()
ens2 = NucleusEnsemble.from_device_layout(CODE_ISOTOPES, 18, decay=True, dt_decay=1e3, layout_seed=2)
ens2.step(2)
for n, scheme in ((1500, "symmetric"), (2600, "symmetric"), (1500, "ordered")):
    pos, isp = make_cloud(n, seed=n)
    c = NucleonCloud(pos, isp, keep_force=True, scheme=scheme)
    c.step(2)
pos, isp = make_cloud(900, seed=5, density=1 / 4)
NucleonCloud(pos, isp, scheme="symmetric").step(1)
zn = torch.full((5000,), (6 << 16) | 8, dtype=torch.int32)
pop = DecayPopulation(zn, dt_decay=180825048000.0 * 0.1, seed=1, watch=((6, 8),))
pop.step(4)
torch.cuda.synchronize()
print("sanitize_small ok", int(ens.event_count.item()), int(ens2.event_count.item()))
