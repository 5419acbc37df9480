#!/bin/bash
# ncu captures of round 2 (run through gpurun, one GPU): each command first runs plain, then under ncu.
TAG=${1:-r02d}
set -x
A="--workload decay --no-extras --no-cpu --steps 2 --warmup 1"
python bench.py $A > gpurun_out/plain_decay_$TAG.log 2>&1 &&
timeout 280 ncu --set full --clock-control none --import-source on -k regex:population -s 1 -c 1 -f \
    -o gpurun_out/prof_population_$TAG python bench.py $A > gpurun_out/ncu_decay_$TAG.log 2>&1
C="--workload ensemble --isotope 92,146 --no-extras --no-cpu --no-e2e --steps 2 --warmup 1"
python bench.py $C > gpurun_out/plain_ens_$TAG.log 2>&1 &&
timeout 280 ncu --set full --clock-control none --import-source on -k regex:ensemble_ring -s 1 -c 1 -f \
    -o gpurun_out/prof_ensemble_$TAG python bench.py $C > gpurun_out/ncu_ens_$TAG.log 2>&1
ls -la gpurun_out/*$TAG.ncu-rep
