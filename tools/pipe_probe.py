#!/usr/bin/env python
"""spmvb200_spmv_host on cfg4 (pinned buffers): per-call time with the kernel chunks in the pipeline and with copies only
(SPMVB200_PIPE_NO_KERNEL=1), for several chunk counts."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import spmv_openmp_cuda_b200 as sp  # noqa: E402
from spmv_openmp_cuda_b200 import synth  # noqa: E402

d = synth.device_csr(synth.banded(1 << 25, 32, 1 << 15))
x = torch.empty(d.N, dtype=torch.float64).pin_memory()
y = torch.empty(d.M, dtype=torch.float64).pin_memory()
x.numpy()[:] = synth.host_vector(d.N)
for _ in range(8):
    sp.spmv_host(sp.CSR_ROWS, d, x, y)
ts = []
for _ in range(20):
    t = time.perf_counter()
    sp.spmv_host(sp.CSR_ROWS, d, x, y)
    ts.append((time.perf_counter() - t) * 1e3)
ts.sort()
print("chunks=%s no_kernel=%s: median %.3f ms  min %.3f ms" % (os.environ.get("SPMVB200_HOST_CHUNKS", "default"), bool(os.environ.get("SPMVB200_PIPE_NO_KERNEL")), ts[10], ts[0]))
