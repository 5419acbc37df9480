#!/bin/bash
# A/B of cloud variants (run through gpurun): each argument is "ENV=VAL,--flag value" (env part optional)
N=${CLOUD_N:-524288}
for cfg in "$@"; do
  envp=$(echo "$cfg" | cut -d, -f1); flags=$(echo "$cfg" | cut -s -d, -f2-)
  env $envp python bench.py --workload cloud --cloud-n $N --steps 3 --warmup 1 --no-extras $flags 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$cfg', 'N', $N, 'pairs/s %.4g' % d['value'], 'ms %.2f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
    elif 'Error' in l or 'error' in l: print(l.rstrip())
"
done
