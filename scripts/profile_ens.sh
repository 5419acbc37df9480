#!/bin/bash
TAG=${1:-x}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ensemble -s 1 -c 1 \
    -f -o gpurun_out/prof_ensemble_${TAG} python bench.py --steps 2 --warmup 1 --no-extras \
    > gpurun_out/ncu_ens_${TAG}.log 2>&1
ls -la gpurun_out | tail -3
