"""C2 host-buffer leg: chunk-count sweep of pyqmd_ensemble_step_host (run on the GPU box)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyqmd_b200.state import HostEnsembleRunner, NucleusEnsemble
ens = NucleusEnsemble.from_templates(((82, 126),), 65536, decay=False)
pairs = ens.pairs_per_step()
out = {}
for ch in (4, 8, 12, 16, 24, 32, 48):
    r = HostEnsembleRunner(ens, chunks=ch)
    for _ in range(3):
        r.step(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        r.step(1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out[ch] = {"ms": round(ms, 3), "pairs_per_s": pairs / ms * 1e3}
    del r
print(json.dumps(out))
