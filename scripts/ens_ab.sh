#!/bin/bash
# A/B of ensemble kernel builds (run through gpurun): args = library tags (pyqmd_b200/libpyqmd_v<tag>.so)
for tag in "$@"; do
  PYQMD_B200_LIB=$PWD/pyqmd_b200/libpyqmd_v$tag.so python bench.py --steps 60 --warmup 5 --no-extras 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$tag', 'pairs/s %.4g' % d['value'], 'ms %.3f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'])
    elif 'rror' in l: print(l.rstrip())
"
done
