#!/bin/bash
# 8-GPU box: bit-identity tests at 2/4/8 ranks, the DEFAULT bench line at 8 (and the headline alone at 4),
# the reference arm as the driver launches it, and the all-ranks PCIe probe.
set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 \
    bench.py --gpus 4 --steps 20 --warmup 5 --no-extras > gpurun_out/bench_final_g4.json 2> gpurun_out/bench_final_g4.err; echo rc=$?
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_final_g8.json 2> gpurun_out/bench_final_g8.err; echo rc=$?
tail -3 gpurun_out/bench_final_g8.err
python - <<PY
import json
for N in (4, 8):
    d = json.loads([l for l in open(f"gpurun_out/bench_final_g{N}.json") if l.startswith("{")][-1])
    print("N=%d value %.4g ms %.3f e2e %.4g exchange %s" % (N, d["value"], d["ms_per_step"], d["e2e"]["value"], d["details"]["exchange_used"]))
    print({k: d["details"].get(k) for k in ("pair_kernel_ms_per_rank", "pair_kernel_imbalance", "step_ms_outside_pair_kernel")})
    print(json.dumps(d["summary"]))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
    scripts/pcie_probe.py > gpurun_out/pcie_probe_g8.json 2> gpurun_out/pcie_probe_g8.err; tail -c 1500 gpurun_out/pcie_probe_g8.json
timeout 200 python scripts/pcie_probe.py > gpurun_out/pcie_probe_g1.json 2>/dev/null; tail -c 600 gpurun_out/pcie_probe_g1.json
