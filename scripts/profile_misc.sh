#!/bin/bash
# ncu captures of the remaining kernels (run through gpurun): population, layout, overlaps, integrate/exchange
TAG=${1:-x}
mkdir -p gpurun_out
cat > /tmp/misc_prof.py <<'PY'
import sys
sys.path.insert(0, '.')
import torch
from pyqmd_b200.state import DecayPopulation, NucleusEnsemble
zn = torch.full((20_000_000,), (6 << 16) | 8, dtype=torch.int32)
pop = DecayPopulation(zn, dt_decay=180825048000.0 * 1e-3, seed=1, watch=((6, 8),))
pop.step(1); pop.step(1)
ens = NucleusEnsemble.from_device_layout(((82, 126),), 16384, decay=False, layout_seed=1)
for _ in range(6):
    ens.frame(4)
torch.cuda.synchronize()
PY
python /tmp/misc_prof.py && ncu --set full --clock-control none --import-source on \
  -k "regex:population_kernel|init_layout_kernel|resolve_overlaps_kernel" -c 6 -f -o gpurun_out/prof_misc_${TAG} \
  python /tmp/misc_prof.py > gpurun_out/ncu_misc_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_misc_${TAG}.log
