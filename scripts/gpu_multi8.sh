#!/bin/bash
# 8-GPU box: bit-identity tests at 2/4/8 ranks, then the bench line at 4 and 8 ranks (and C5 / C3 at 8).
set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -5
for N in 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N --steps 10 --warmup 3 --no-extras > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err; echo rc=$?
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_g$N.json") if l.startswith("{")][-1])
print("N=$N value %.4g ms %.3f e2e %.4g exchange %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["details"]["exchange_used"]))
print({k: d["details"].get(k) for k in ("pair_kernel_ms_per_rank", "pair_kernel_imbalance", "step_ms_outside_pair_kernel")})
PY
done
for W in decay mixed ensemble; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
    bench.py --gpus 8 --workload $W --steps 10 --warmup 3 --no-extras > gpurun_out/bench_${W}_g8.json 2> gpurun_out/bench_${W}_g8.err; echo rc=$?
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_${W}_g8.json") if l.startswith("{")][-1])
print("$W N=8 value %.4g ms %.4f e2e %s" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value")))
PY
done
