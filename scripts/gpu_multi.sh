#!/bin/bash
# Multi-GPU check (run through gpurun --gpus N): bit-identity tests + the bench line at N ranks.
N=${1:-2}
set -x
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-extras > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err; echo rc=$?
tail -3 gpurun_out/bench_g$N.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_g$N.json") if l.startswith("{")][-1])
print("value %.4g ms %.3f e2e %.4g exchange %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["exchange"]))
print({k: d["config"].get(k) for k in ("pair_kernel_ms_per_rank", "pair_kernel_imbalance", "step_ms_outside_pair_kernel")})
PY
