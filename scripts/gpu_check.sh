#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_forces.py tests/test_gpu_overlaps.py tests/test_gpu_decay.py tests/test_gpu_fuzz.py -q 2>&1 | tail -5
for args in "--workload ensemble" "--workload ensemble --settled" "--workload ensemble --settled --list-order" "--workload ensemble --isotope 92,146 --settled" "--workload ensemble --isotope 92,146 --settled --list-order" "--workload ensemble --isotope 92,146" "--workload mixed"; do
timeout 200 python bench.py $args --no-extras --no-cpu --no-e2e --steps 20 --warmup 5 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$args', '%.4g' % d['value'], 'ms', '%.3f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], 'flops/pair %.2f' % d['roofline']['flops_per_pair'])"
done
