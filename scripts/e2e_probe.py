import sys, json
sys.path.insert(0, '.')
import torch
from pyqmd_b200.state import NucleusEnsemble, HostEnsembleRunner
ens = NucleusEnsemble.from_templates(((82,126),), 65536, decay=False)
def t(fn, k):
    fn(); torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/k
out={}
h=torch.empty(232*1024*1024, dtype=torch.uint8, pin_memory=True); d=torch.empty_like(h, device='cuda')
out['h2d_GBs']=h.numel()/t(lambda: d.copy_(h, non_blocking=True),5)/1e6
out['d2h_GBs']=h.numel()/t(lambda: h.copy_(d, non_blocking=True),5)/1e6
for chunks, streams in ((4,3),(8,3),(16,3),(24,3),(32,3),(64,3)):
    r=HostEnsembleRunner(ens, chunks=chunks)
    out[f'e2e_ms_c{chunks}_s{streams}']=t(lambda: r.step(1),8)
print(json.dumps(out))
