#!/bin/bash
# cloud symmetric kernel: library variants x tiles-per-unit (run through gpurun)
N=${CLOUD_N:-1000000}
run() { env $1 python bench.py --workload cloud --cloud-n $N --steps 3 --warmup 1 --no-extras 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$1', 'pairs/s %.4g' % d['value'], 'ms %.2f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'])
    elif 'rror' in l: print(l.rstrip()[:200])
"; }
run "A=1"
for tag in U2 U4; do run "PYQMD_B200_LIB=$PWD/pyqmd_b200/libpyqmd_v$tag.so"; done
