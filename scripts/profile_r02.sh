#!/bin/bash
# ncu captures of round 2, final kernels (run through gpurun, one GPU): plain run first, then under ncu.
TAG=${1:-r02g}
set -x
timeout 200 bash scripts/lib_ab.sh "--workload ensemble --isotope 92,146 --no-cpu --no-e2e --steps 40 --warmup 5" - p9 p10
A="--workload decay --no-extras --no-cpu --steps 2 --warmup 1"
python bench.py $A > gpurun_out/plain_decay_$TAG.log 2>&1 &&
timeout 280 ncu --set full --clock-control none --import-source on -k regex:population -s 1 -c 1 -f \
    -o gpurun_out/prof_population_$TAG python bench.py $A > gpurun_out/ncu_decay_$TAG.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches_$TAG.log 2>&1
ls -la gpurun_out/*$TAG*
