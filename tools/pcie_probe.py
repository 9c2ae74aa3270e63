#!/usr/bin/env python
"""PCIe probe: H2D, D2H and both at once (pinned buffers, two streams) for the e2e sizes."""
import torch, time
n = 2097152
hx = torch.empty(n, dtype=torch.float64).pin_memory(); hy = torch.empty(n, dtype=torch.float64).pin_memory()
dx = torch.empty(n, dtype=torch.float64, device="cuda"); dy = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(f, reps=50):
    for _ in range(5): f()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / reps * 1e3
def h2d():
    with torch.cuda.stream(s1): dx.copy_(hx, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): hy.copy_(dy, non_blocking=True)
def both(): h2d(); d2h()
def chunks(k=8):
    c = n // k
    for i in range(k):
        with torch.cuda.stream(s1): dx[i*c:(i+1)*c].copy_(hx[i*c:(i+1)*c], non_blocking=True)
        with torch.cuda.stream(s2): hy[i*c:(i+1)*c].copy_(dy[i*c:(i+1)*c], non_blocking=True)
for name, f in (("h2d", h2d), ("d2h", d2h), ("both (2 streams)", both), ("both, 8 chunks each", chunks)):
    ms = run(f); print("%-22s %.3f ms  -> %.1f GB/s per direction" % (name, ms, n * 8 / ms / 1e6))
