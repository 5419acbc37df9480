import sys, time, json
sys.path.insert(0, '.')
import torch
from pyqmd_b200.state import NucleusEnsemble, README_ISOTOPES
out = {}
for name, iso, n in (("pb208_65536", ((82, 126),), 65536), ("mixed_1M", README_ISOTOPES, 1_000_000)):
    ens = NucleusEnsemble.from_device_layout(iso, n, decay=False, layout_seed=1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ens.init_layout(2)
    torch.cuda.synchronize()
    out[name] = {"seconds": time.perf_counter() - t0, "nuclei": n}
    del ens
print(json.dumps(out))
