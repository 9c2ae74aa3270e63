"""Host-buffer path (spmvb200_spmv_host) on cfg2: wall time per call for every (direct-y, chunks) combination.

direct-y = the kernels store y straight into the caller's page-locked buffer (SPMVB200_HOST_DIRECT_Y: 1 = default,
2 = every single launch, 3 = chunked pipeline too; the r01d log was taken when 1 meant "chunked only" and 2 "everything");
0 = y goes to device memory and comes down as copy-engine jobs.  Every variant's y is compared with the device-path result.
    python tools/e2e_probe.py [--dbg]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import spmv_openmp_cuda_b200 as sp
from spmv_openmp_cuda_b200 import synth

d = synth.device_csr(synth.stencil27(128)); dm = d.to_ell(sp.FMT_ELL_COLMAJOR)
hx = torch.empty(dm.N, dtype=torch.float64).pin_memory(); hy = torch.empty(dm.M, dtype=torch.float64).pin_memory()
hx.copy_(torch.from_numpy(synth.host_vector(dm.N)))
dx = sp.DeviceVector.from_host(hx.numpy()); dy = sp.DeviceVector(dm.M)
sp.cudaSpMVRowsELL(dm, dx, sp.Config(), dy)
y_ref = dy.to_host()
KINDS = (("ELL_ROWS", sp.ELL_ROWS, dm), ("CSR_ROWS", sp.CSR_ROWS, d), ("CSR_ROWS_WARP", sp.CSR_ROWS_WARP, d))
for name, kind, m in (KINDS[:1] if "--ell-only" in sys.argv else KINDS):
    for direct in (0, 1, 2, 3):
        for chunks in (1, 2, 4, 8, 16):
            os.environ["SPMVB200_HOST_DIRECT_Y"] = str(direct)
            os.environ["SPMVB200_HOST_CHUNKS"] = str(chunks)
            for i in range(5): sp.spmv_host(kind, m, hx, hy)
            if "--dbg" in sys.argv and chunks == 4:
                os.environ["SPMVB200_PIPE_DEBUG"] = "1"; sp.spmv_host(kind, m, hx, hy); del os.environ["SPMVB200_PIPE_DEBUG"]
            hy.fill_(float("nan"))
            torch.cuda.synchronize()
            t = time.perf_counter()
            for i in range(100): ms = sp.spmv_host(kind, m, hx, hy)
            wall = (time.perf_counter() - t) * 10
            y = hy.numpy()
            ok = np.array_equal(y, y_ref) if kind != sp.CSR_ROWS_WARP else bool(np.allclose(y, y_ref, rtol=1e-12, atol=1e-18))
            print("%-14s direct_y=%d chunks=%-2d wall per call %.3f ms  kernel %.3f ms  %s" % (name, direct, chunks, wall, ms, "ok" if ok else "MISMATCH"), flush=True)
