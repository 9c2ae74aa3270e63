import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import spmv_openmp_cuda_b200 as sp
from spmv_openmp_cuda_b200 import synth, capi
d = synth.device_csr(synth.stencil27(128)); dm = d.to_ell(sp.FMT_ELL_COLMAJOR)
hx = torch.empty(dm.N, dtype=torch.float64).pin_memory(); hy = torch.empty(dm.M, dtype=torch.float64).pin_memory()
hx.copy_(torch.from_numpy(synth.host_vector(dm.N)))
for kind, m in ((sp.ELL_ROWS, dm), (sp.CSR_ROWS, d)):
    for i in range(3): sp.spmv_host(kind, m, hx, hy)
    if os.environ.get("DBG"):
        os.environ["SPMVB200_PIPE_DEBUG"] = "1"; sp.spmv_host(kind, m, hx, hy); del os.environ["SPMVB200_PIPE_DEBUG"]
    t = time.perf_counter()
    for i in range(100): ms = sp.spmv_host(kind, m, hx, hy)
    print("kind", kind, "wall per call %.3f ms, kernel %.3f ms" % ((time.perf_counter() - t) * 10, ms))
