#!/bin/bash
# One GPU call: smoke, the -m gpu suite and the default bench line (run through gpurun).
set -x
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final_g1.json 2> gpurun_out/bench_final_g1.err; echo rc=$?
tail -3 gpurun_out/bench_final_g1.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_final_g1.json") if l.startswith("{")][-1])
print(json.dumps(d["summary"]))
print(json.dumps(d["also"]["c1"])[:1200])
PY
C="--workload ensemble --isotope 92,146 --nuclei 1 --no-extras --no-cpu --no-e2e --steps 3 --warmup 1 --substeps 100"
python bench.py $C > gpurun_out/plain_c1_r02i.log 2>&1 &&
timeout 280 ncu --set full --clock-control none --import-source on -k regex:ensemble_cluster -s 1 -c 1 -f \
    -o gpurun_out/prof_ensemble_r02i python bench.py $C > gpurun_out/ncu_c1_r02i.log 2>&1
ls -la gpurun_out/prof_ensemble_r02i.ncu-rep
