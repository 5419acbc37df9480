#!/bin/bash
# ncu capture of the cloud force kernel (run through gpurun): bash scripts/profile_cloud.sh <tag> [scheme]
TAG=${1:-x}; SCHEME=${2:-symmetric}
mkdir -p gpurun_out
python bench.py --workload cloud --cloud-n 262144 --steps 1 --warmup 1 --no-extras --cloud-scheme $SCHEME > gpurun_out/plain3_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cloud_sym_kernel\|cloud_force_kernel -s 1 -c 1 \
    -f -o gpurun_out/prof_cloud_${TAG} python bench.py --workload cloud --cloud-n 262144 --steps 1 --warmup 1 --no-extras --cloud-scheme $SCHEME \
    > gpurun_out/ncu_cloud_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_cloud_${TAG}.log
ls -la gpurun_out | grep ${TAG}
