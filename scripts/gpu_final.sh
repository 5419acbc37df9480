#!/bin/bash
# Round-end evidence on one GPU: default bench line, reference arm, ncu launch list of the same command.
TAG=${1:-r02c}
set -x
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo rc=$?
tail -c 1500 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo rc=$?
tail -c 600 gpurun_out/bench_ref_$TAG.json
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches_$TAG.log 2>&1
tail -2 gpurun_out/ncu_launches_$TAG.log | cut -c1-300
