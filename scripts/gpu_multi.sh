#!/bin/bash
# Multi-GPU check (run through gpurun --gpus N): bit-identity tests + the DEFAULT bench line at N ranks.
N=${1:-2}
set -x
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_full_g$N.json 2> gpurun_out/bench_full_g$N.err; echo rc=$?
tail -3 gpurun_out/bench_full_g$N.err
tail -c 1200 gpurun_out/bench_full_g$N.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 3 --warmup 1 | tail -c 400
