#!/bin/bash
# One GPU call: smoke and the -m gpu suite (run through gpurun).
set -x
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25
for k in cluster block ring; do
PYQMD_ENSEMBLE_KERNEL=$k timeout 120 python bench.py --workload ensemble --isotope 92,146 --nuclei 1 --no-extras --no-cpu --no-e2e --steps 100 --warmup 10 --substeps 100 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$k: one U-238, 100 fused sub-steps: us per sub-step', d['ms_per_step'] * 10)"
done
