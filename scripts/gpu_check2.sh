#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_cloud.py -q 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err; echo rc=$?
tail -3 gpurun_out/bench_r02f.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_r02f.json") if l.startswith("{")][-1])
print(json.dumps(d["summary"], indent=0))
print(d["also"]["cloud_c4_skip_exact_zeros_optin"].get("bit_identical_to_default_after_1_step"), d["also"]["cloud_c4_skip_exact_zeros_optin"].get("ms_per_step"))
PY
