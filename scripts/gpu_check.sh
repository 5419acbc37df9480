#!/bin/bash
# One GPU call: smoke, the -m gpu suite, the default bench line and the reference arm (run through gpurun).
set -x
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo rc=$?
tail -c 1800 gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?
tail -c 900 gpurun_out/bench_ref.json
