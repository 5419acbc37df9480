#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_decay.py tests/test_gpu_sim.py -q 2>&1 | tail -5
for tag in - pm3 pm4; do
if [ "$tag" = "-" ]; then LIB=""; else LIB=$PWD/pyqmd_b200/libpyqmd_v$tag.so; fi
PYQMD_B200_LIB=$LIB timeout 200 python bench.py --workload decay --no-extras --no-cpu --steps 20 --warmup 5 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$tag decay', '%.4g' % d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'])"
done
timeout 200 python bench.py --workload decay --substeps 100 --no-extras --no-cpu --steps 5 --warmup 2 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('decay x100', '%.4g' % d['value'], 'ms', d['ms_per_step'])"
