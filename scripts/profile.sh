#!/bin/bash
# Profiling pass on the GPU box (run through gpurun):  bash scripts/profile.sh <tag>
# 1. launch list of the default bench command (share of each kernel in the step)
# 2. one full ncu capture of each hot kernel
TAG=${1:-r01}
mkdir -p gpurun_out
set -x
python bench.py --steps 3 --warmup 1 --no-extras > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 3 --warmup 1 --no-extras \
    > gpurun_out/ncu_launches_${TAG}.log 2>&1
python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ensemble -s 1 -c 1 \
    -f -o gpurun_out/prof_ensemble_${TAG} python bench.py --steps 2 --warmup 1 --no-extras \
    > gpurun_out/ncu_ens_${TAG}.log 2>&1
python bench.py --workload cloud --cloud-n 262144 --steps 1 --warmup 1 > gpurun_out/plain3_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:cloud_sym_kernel|cloud_force_kernel" -s 1 -c 1 \
    -f -o gpurun_out/prof_cloud_${TAG} python bench.py --workload cloud --cloud-n 262144 --steps 1 --warmup 1 \
    > gpurun_out/ncu_cloud_${TAG}.log 2>&1
ls -la gpurun_out
