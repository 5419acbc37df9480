#!/bin/bash
# Round-end evidence on one GPU (run through gpurun): smoke, the -m gpu suite, the default bench line, the
# reference arm, then the ncu launch list and one --set full capture of the dominant kernel of the same command.
TAG=${1:-r02k}
set -x
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 1500 python -m pytest tests -m gpu -q ) 2>&1 | tail -8
( time timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ); echo rc=$?
tail -c 1500 gpurun_out/bench_$TAG.json
( time timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err ); echo rc=$?
tail -c 700 gpurun_out/bench_ref_$TAG.json
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches_$TAG.log 2>&1
C="--steps 2 --warmup 1 --no-cpu --no-e2e --no-extras"
python bench.py $C > gpurun_out/plain_cloud_$TAG.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cloud_sym_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_cloud_$TAG python bench.py $C > gpurun_out/ncu_cloud_$TAG.log 2>&1
ls -la gpurun_out/*$TAG*
