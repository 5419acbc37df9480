"""Multi-GPU check of the cloud schemes (run under torchrun through gpurun --gpus N):
every rank steps the same N-nucleon cloud sharded over the world, rank 0 also steps a single-GPU
instance; the symmetric scheme must agree BIT FOR BIT (integer force accumulation), the ordered
scheme to FP32 rounding."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from bench import make_cloud
from pyqmd_b200.state import NucleonCloud

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 150_001
pos, isp = make_cloud(n, seed=99)
out = {"n": n, "world": world}
for scheme, exchange in (("symmetric", "peer"), ("symmetric", "nccl"), ("ordered", "nccl")):
    multi = NucleonCloud(pos, isp, device=f"cuda:{local}", rank=rank, world=world, scheme=scheme,
                         exchange=exchange)
    used = multi.exchange
    scheme = scheme if scheme == "ordered" else f"symmetric_{exchange}"
    multi.step(3)
    torch.cuda.synchronize()
    if rank == 0:
        single = NucleonCloud(pos, isp, device="cuda:0", scheme=scheme.split("_")[0])
        single.step(3)
        a, b = multi.pos[:n], single.pos[:n]
        out[scheme] = {"bit_identical": bool(torch.equal(a, b)),
                       "max_abs_diff": float((a - b).abs().max()), "exchange_used": used}
    # replicas must be identical on every rank
    ref = multi.pos.clone()
    dist.broadcast(ref, 0)
    same = torch.tensor([int(torch.equal(ref, multi.pos))], device=f"cuda:{local}")
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    if rank == 0:
        out[scheme]["replicas_identical"] = bool(same.item())
# host-resident blocks (NucleonCloud.step_host): upload block -> all-gather -> step -> download block must
# reproduce the device-resident steps bit for bit
multi = NucleonCloud(pos, isp, device=f"cuda:{local}", rank=rank, world=world)
blk = multi.i1 - multi.i0
h_pos = torch.empty(blk, 2, dtype=torch.float32).pin_memory()
h_vel = torch.empty(blk, 2, dtype=torch.float32).pin_memory()
multi.download_block(h_pos, h_vel)
for _ in range(3):
    multi.step_host(h_pos, h_vel)
if rank == 0:
    single = NucleonCloud(pos, isp, device="cuda:0")
    single.step(3)
    out["host_step"] = {"bit_identical": bool(torch.equal(multi.pos[:n], single.pos[:n])
                                              and torch.equal(h_pos.cuda(), single.pos[multi.i0:multi.i1])),
                        "exchange_used": multi.exchange}
if rank == 0:
    print(json.dumps(out))
    for k in ("symmetric_peer", "symmetric_nccl"):
        assert out[k]["bit_identical"] and out[k]["replicas_identical"], k
    assert out["ordered"]["max_abs_diff"] < 1e-3 and out["ordered"]["replicas_identical"]
dist.destroy_process_group()
