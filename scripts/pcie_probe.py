"""PCIe probe: H2D alone, D2H alone, both at once on two streams (pinned memory)."""
import json
import torch
n = 232 * 1024 * 1024
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, k=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
def both(chunks):
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    c = n // chunks
    for i in range(chunks):
        with torch.cuda.stream(s1): d_in[i*c:(i+1)*c].copy_(h_in[i*c:(i+1)*c], non_blocking=True)
        with torch.cuda.stream(s2): h_out[i*c:(i+1)*c].copy_(d_out[i*c:(i+1)*c], non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)
out = {"h2d_ms": timed(lambda: d_in.copy_(h_in, non_blocking=True)), "d2h_ms": timed(lambda: h_out.copy_(d_out, non_blocking=True))}
for ch in (1, 8, 32):
    out[f"both_ms_chunks{ch}"] = timed(lambda: both(ch))
out["GBs"] = {k: n / v / 1e6 * (2 if k.startswith("both") else 1) for k, v in out.items() if k.endswith("ms") or "ms_" in k}
print(json.dumps(out))
