"""PCIe probe: H2D alone, D2H alone, both at once on two streams (pinned memory).

Single process:            python scripts/pcie_probe.py
All GPUs of a box at once: python -m torch.distributed.run --nproc-per-node 8 scripts/pcie_probe.py
(every rank copies 232 MB each way at the same time, the pattern of the C2 host-buffer leg of bench.py;
rank 0 prints per-rank and aggregate GB/s -- the host ceiling as a number)."""
import json
import os

import torch

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 232 * 1024 * 1024
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, k=5):
    fn(); torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


def both(chunks):
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    c = n // chunks
    for i in range(chunks):
        with torch.cuda.stream(s1):
            d_in[i * c:(i + 1) * c].copy_(h_in[i * c:(i + 1) * c], non_blocking=True)
        with torch.cuda.stream(s2):
            h_out[i * c:(i + 1) * c].copy_(d_out[i * c:(i + 1) * c], non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)


out = {"h2d_ms": timed(lambda: d_in.copy_(h_in, non_blocking=True)),
       "d2h_ms": timed(lambda: h_out.copy_(d_out, non_blocking=True))}
for ch in (1, 8, 32):
    out[f"both_ms_chunks{ch}"] = timed(lambda: both(ch))
gbs = {k: n / v / 1e6 * (2 if k.startswith("both") else 1) for k, v in out.items()}
if dist is not None:
    keys = sorted(gbs)
    t = torch.tensor([gbs[k] for k in keys], device="cuda", dtype=torch.float64)
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    if rank == 0:
        per_rank = {k: [round(float(a[i]), 1) for a in allt] for i, k in enumerate(keys)}
        print(json.dumps({"world": world, "bytes_each_way": n, "GBs_per_rank": per_rank,
                          "GBs_aggregate": {k: round(sum(v), 1) for k, v in per_rank.items()}}))
    dist.destroy_process_group()
else:
    out["GBs"] = gbs
    print(json.dumps(out))
