"""Multi-GPU parity of the cloud schemes (skipped on boxes with fewer than 2 GPUs): launches
scripts/check_cloud_multi.py under torchrun -- symmetric scheme with the fused peer-memory exchange and
with the NCCL exchange must be BIT-IDENTICAL to the single-GPU run, replicas identical on all ranks."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_cloud_schemes_bit_identical_across_gpus(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "scripts", "check_cloud_multi.py"), "60001"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == world
    for k in ("symmetric_peer", "symmetric_nccl", "ordered"):
        assert out[k]["replicas_identical"], k
    assert out["symmetric_peer"]["bit_identical"] and out["symmetric_nccl"]["bit_identical"]
    # the fused peer-memory kernel really ran (the NCCL fallback is opt-in and was not requested)
    assert out["symmetric_peer"]["exchange_used"] == "peer"
    assert out["symmetric_nccl"]["exchange_used"] == "nccl"
    assert out["host_step"]["bit_identical"], out["host_step"]


def test_one_process_two_devices():
    """Per-device kernel attributes and device-pinned C-ABI calls: an ensemble with the large
    shared-memory configuration on cuda:0, then the same on cuda:1 from the same process, while cuda:0
    stays the current device; a cloud and a decay population on cuda:1 as well."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import numpy as np

    sys.path.insert(0, ROOT)
    from bench import make_cloud
    from pyqmd_b200.state import DecayPopulation, NucleonCloud, NucleusEnsemble
    torch.cuda.set_device(0)
    out = []
    for dev in ("cuda:0", "cuda:1"):
        ens = NucleusEnsemble.from_templates(((92, 146), (82, 126), (6, 8)), 600, device=dev, decay=False)
        ens.step(3)
        ens.frame(2)
        out.append(ens.pos.cpu().numpy())
    assert torch.cuda.current_device() == 0
    assert np.array_equal(out[0], out[1])
    pos, isp = make_cloud(5000, seed=3)
    a = NucleonCloud(pos, isp, device="cuda:0"); b = NucleonCloud(pos, isp, device="cuda:1")
    a.step(2); b.step(2)
    assert torch.equal(a.pos.cpu(), b.pos.cpu())
    zn = torch.full((4096,), (6 << 16) | 8, dtype=torch.int32)
    pa = DecayPopulation(zn, device="cuda:0", dt_decay=1e10, seed=3)
    pb = DecayPopulation(zn, device="cuda:1", dt_decay=1e10, seed=3)
    ca, _ = pa.step(4); cb, _ = pb.step(4)
    assert torch.equal(ca.cpu(), cb.cpu()) and int(ca.sum()) > 0
