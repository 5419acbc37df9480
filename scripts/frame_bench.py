import sys, time, json
sys.path.insert(0, '.')
import torch
from pyqmd_b200.state import NucleusEnsemble
ens = NucleusEnsemble.from_templates(((82,126),), 65536, decay=False)
def t(fn, k):
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/k
out={}
out['proj_fresh_ms']=t(ens.resolve_overlaps,1)
for _ in range(10): ens.frame(4)
out['proj_spread_ms']=t(ens.resolve_overlaps,5)
out['step4_spread_ms']=t(lambda: ens.step(4),5)
out['step1_spread_ms']=t(lambda: ens.step(1),20)
out['frame_ms']=t(lambda: ens.frame(4),10)
pos=ens.pos.view(65536,208,2); ext=(pos-pos.mean(1,keepdim=True)).norm(dim=2).max(1).values
out['extent_mean']=float(ext.mean())
print(json.dumps(out))
