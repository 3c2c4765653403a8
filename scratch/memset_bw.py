# pure-write roof: cudaMemset / torch fill of a 5 GB buffer (what the similarity kernel's store stream competes with)
import json, torch
x = torch.empty(5 * 1024**3 // 4, dtype=torch.float32, device="cuda")
for name, fn in (("memset_zero", lambda: x.zero_()), ("fill_1.5", lambda: x.fill_(1.5))):
    ts = []
    for _ in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print(json.dumps(dict(op=name, ms=round(min(ts), 4), gbs=round(x.numel() * 4 / min(ts) / 1e6, 1))), flush=True)
