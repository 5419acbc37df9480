#!/bin/bash
# A/B of the ensemble kernels on one box: PYQMD_ENSEMBLE_KERNEL pins the kernel (run through gpurun).
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py $ARGS --steps 10 --warmup 3 --no-cpu --no-e2e --no-extras 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$name', 'value %.4g' % d['value'], 'ms %.3f' % d['ms_per_step'])
    elif 'rror' in l: print(l.rstrip()[:300])
"
}
ARGS="--workload ensemble --settled --substeps 4"
run pb_settled4_quad PYQMD_ENSEMBLE_KERNEL=quad
run pb_settled4_block PYQMD_ENSEMBLE_KERNEL=block
ARGS="--workload ensemble --settled --substeps 1"
run pb_settled1_quad PYQMD_ENSEMBLE_KERNEL=quad
run pb_settled1_block PYQMD_ENSEMBLE_KERNEL=block
ARGS="--workload ensemble --substeps 4"
run pb_free4_quad PYQMD_ENSEMBLE_KERNEL=quad
run pb_free4_block PYQMD_ENSEMBLE_KERNEL=block
