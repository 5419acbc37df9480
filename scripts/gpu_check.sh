#!/bin/bash
# One GPU call: smoke, the -m gpu suite and the default bench line (run through gpurun).
set -x
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final_g1.json 2> gpurun_out/bench_final_g1.err; echo rc=$?
tail -3 gpurun_out/bench_final_g1.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_final_g1.json") if l.startswith("{")][-1])
print(json.dumps(d["summary"]))
print(d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["config"], list(d.keys()))
PY
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 | tail -c 700
