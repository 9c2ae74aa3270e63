#!/usr/bin/env python
"""spmvb200_spmv_host on cfg4 (pinned buffers): chunk schedule (uniform / ramp) x chunk count x copy streams per direction,
with and without the kernel chunks (SPMVB200_PIPE_NO_KERNEL is read once per process: pass --no-kernel to set it)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if "--no-kernel" in sys.argv:
    os.environ["SPMVB200_PIPE_NO_KERNEL"] = "1"
import torch  # noqa: E402

TIMED = "--timed" in sys.argv

import spmv_openmp_cuda_b200 as sp  # noqa: E402
from spmv_openmp_cuda_b200 import synth  # noqa: E402

d = synth.device_csr(synth.banded(1 << 25, 32, 1 << 15))
x = torch.empty(d.N, dtype=torch.float64).pin_memory()
y = torch.empty(d.M, dtype=torch.float64).pin_memory()
x.numpy()[:] = synth.host_vector(d.N)
os.environ["SPMVB200_HOST_CHUNKS"] = "4"
for _ in range(8):
    sp.spmv_host(sp.CSR_ROWS, d, x, y, timed=TIMED)
prev = 4
for streams in ("1",):
    for sched in ("uniform", "ramp"):
        for chunks in (12, 16, 20, 24, 12, 32):
            if chunks == prev:
                continue
            prev = chunks
            os.environ["SPMVB200_HOST_COPY_STREAMS"] = streams
            os.environ["SPMVB200_HOST_SCHED"] = sched
            os.environ["SPMVB200_HOST_CHUNKS"] = str(chunks)
            for _ in range(4):
                sp.spmv_host(sp.CSR_ROWS, d, x, y, timed=TIMED)
            ts = []
            for _ in range(20):
                t = time.perf_counter()
                sp.spmv_host(sp.CSR_ROWS, d, x, y, timed=TIMED)
                ts.append((time.perf_counter() - t) * 1e3)
            ts.sort()
            print("streams=%s sched=%-7s chunks=%2d no_kernel=%s: median %.3f ms  min %.3f ms" % (
                streams, sched, chunks, bool(os.environ.get("SPMVB200_PIPE_NO_KERNEL")), ts[10], ts[0]), flush=True)
if "--timeline" in sys.argv:
    os.environ["SPMVB200_HOST_COPY_STREAMS"] = "1"
    os.environ["SPMVB200_HOST_SCHED"] = "ramp"
    os.environ["SPMVB200_HOST_CHUNKS"] = "16"
    for _ in range(3):
        sp.spmv_host(sp.CSR_ROWS, d, x, y, timed=TIMED)
    os.environ["SPMVB200_PIPE_DEBUG"] = "1"
    sp.spmv_host(sp.CSR_ROWS, d, x, y, timed=TIMED)
    sp.spmv_host(sp.CSR_ROWS, d, x, y, timed=TIMED)
