#!/usr/bin/env python
"""Why does the chunked host path reach only ~0.8 of the duplex link ceiling?  Bare experiments with pinned buffers of the cfg4 size
(268 MB each way): whole-vector duplex copies, 16-chunk duplex copies, and both again while an HBM-saturating kernel runs."""
import sys
import torch

n = 1 << 25
hx, hy = torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory()
dx, dy = torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty(n, dtype=torch.float64, device="cuda")
big_a, big_b = torch.empty(1 << 28, dtype=torch.float64, device="cuda"), torch.empty(1 << 28, dtype=torch.float64, device="cuda")
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def run(chunks, hog, stagger=False):
    best = 1e9
    for _ in range(6):
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(s1)
        s2.wait_event(e0)
        s3.wait_event(e0)
        if hog:
            with torch.cuda.stream(s3):
                for _ in range(3):
                    big_b.copy_(big_a)  # 2 GB read + 2 GB write each: ~0.7 ms of saturated HBM
        c = n // chunks
        ups = []
        with torch.cuda.stream(s1):
            for k in range(chunks):
                dx[k * c:(k + 1) * c].copy_(hx[k * c:(k + 1) * c], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s1)
                ups.append(ev)
        with torch.cuda.stream(s2):
            for k in range(chunks):
                if stagger:
                    s2.wait_event(ups[k])  # download chunk k only after upload piece k (as the pipeline does)
                hy[k * c:(k + 1) * c].copy_(dy[k * c:(k + 1) * c], non_blocking=True)
        e1.record(s1)
        e2.record(s2)
        torch.cuda.synchronize()
        best = min(best, max(e0.elapsed_time(e1), e0.elapsed_time(e2)))
    return best


for chunks in (1, 4, 16, 64):
    print("chunks %2d  duplex %.3f ms   staggered %.3f ms   with HBM hog %.3f ms   staggered + hog %.3f ms" % (
        chunks, run(chunks, False), run(chunks, False, True), run(chunks, True), run(chunks, True, True)), flush=True)
