#!/bin/bash
# Sweep of the polynomial-2^x share of the cloud far path (run through gpurun).
mkdir -p gpurun_out
python -m pytest tests/test_gpu_cloud.py tests/test_gpu_sim.py -x -q 2>&1 | tail -5
for np in 0 1 2 3 4; do
  PYQMD_CLOUD_NPOLY=$np python bench.py --workload cloud --cloud-n ${1:-524288} --steps 3 --warmup 1 --no-extras 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('npoly', $np, 'pairs/s %.4g' % d['value'], 'ms %.2f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], d['clocks'])
"
done
