"""Development probe: pinned H2D / D2H bandwidth for the e2e sizes (8 MB J in, 8 MB J + 8 MB pi out)."""
import torch, time, json
N = 1001 * 1001
h_in = torch.empty(N, dtype=torch.float64).pin_memory(); d = torch.empty(N, dtype=torch.float64, device="cuda")
h_out = torch.empty(2 * N, dtype=torch.float64).pin_memory(); d2 = torch.empty(2 * N, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, k=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / k
up = t(lambda: d.copy_(h_in, non_blocking=True)); dn = t(lambda: h_out.copy_(d2, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): both()
torch.cuda.synchronize(); bi = (time.perf_counter() - t0) / 20 * 1e3
print(json.dumps({"h2d_8MB_ms": up, "h2d_GBs": 8 * N / up / 1e6, "d2h_16MB_ms": dn, "d2h_GBs": 16 * N / dn / 1e6, "both_directions_ms": bi}))
