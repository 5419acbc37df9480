#!/bin/bash
# 8-GPU box, final build: bit-identity tests at 2/4/8 ranks, the DEFAULT bench line at 8 GPUs as the driver
# launches it, and the headline alone at 1 GPU of the same box (the denominator of the speed-up).
TAG=${1:-r02k}
set -x
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -4
timeout 300 python bench.py --no-extras --no-cpu > gpurun_out/bench_${TAG}_g1box8.json 2> gpurun_out/bench_${TAG}_g1box8.err; echo rc=$?
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus 8 > gpurun_out/bench_${TAG}_g8.json 2> gpurun_out/bench_${TAG}_g8.err; echo rc=$?
tail -3 gpurun_out/bench_${TAG}_g8.err
python - <<PY
import json
one = json.loads([l for l in open("gpurun_out/bench_${TAG}_g1box8.json") if l.startswith("{")][-1])
d = json.loads([l for l in open("gpurun_out/bench_${TAG}_g8.json") if l.startswith("{")][-1])
print("N=1 value %.4g ms %.3f e2e %.4g" % (one["value"], one["ms_per_step"], one["e2e"]["value"]))
print("N=8 value %.4g ms %.3f e2e %.4g exchange %s  speed-up %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["details"]["exchange_used"], d["value"] / one["value"]))
print({k: d["details"].get(k) for k in ("pair_kernel_ms_per_rank", "pair_kernel_imbalance", "step_ms_outside_pair_kernel")})
print(json.dumps(d["summary"]))
PY
