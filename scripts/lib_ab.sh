#!/bin/bash
# A/B of library builds on one bench command (run through gpurun):
#   bash scripts/lib_ab.sh "<bench args>" <tag> [<tag> ...]     tags name pyqmd_b200/libpyqmd_v<tag>.so; "-" = default build
ARGS=$1; shift
for tag in "$@"; do
  if [ "$tag" = "-" ]; then LIB=""; else LIB=$PWD/pyqmd_b200/libpyqmd_v$tag.so; fi
  PYQMD_B200_LIB=$LIB python bench.py $ARGS --no-extras 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$tag', 'value %.4g' % d['value'], 'ms %.3f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'])
    elif 'rror' in l: print(l.rstrip()[:200])
"
done
