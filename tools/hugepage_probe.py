"""Does the physical layout of the caller's page-locked buffers explain the run-to-run spread of the end-to-end time?
Times spmvb200_spmv_host on cfg2 with (a) torch pin_memory buffers (cudaHostAlloc) and (b) 2 MB aligned buffers with
madvise(MADV_HUGEPAGE), touched, then cudaHostRegister'ed."""
import ctypes, mmap, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import spmv_openmp_cuda_b200 as sp
from spmv_openmp_cuda_b200 import synth

print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), flush=True)
libc = ctypes.CDLL("libc.so.6", use_errno=True)
libc.madvise.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
rt = torch.cuda.cudart()


def huge_pinned(n):
    size = ((n * 8 + (2 << 20) - 1) // (2 << 20)) * (2 << 20)
    raw = np.empty(size + (2 << 20), dtype=np.uint8)
    off = (-raw.ctypes.data) % (2 << 20)
    buf = raw[off:off + size]
    rc = libc.madvise(buf.ctypes.data, size, 14)  # MADV_HUGEPAGE
    buf[:] = 0
    err = rt.cudaHostRegister(buf.ctypes.data, size, 0)
    print("madvise rc", rc, "cudaHostRegister", err, flush=True)
    return buf[:n * 8].view(np.float64), raw


torch.cuda.init()
d = synth.device_csr(synth.stencil27(128)); dm = d.to_ell(sp.FMT_ELL_COLMAJOR)
xh = synth.host_vector(dm.N)
tx = torch.empty(dm.N, dtype=torch.float64).pin_memory(); ty = torch.empty(dm.M, dtype=torch.float64).pin_memory()
tx.copy_(torch.from_numpy(xh))
hx, _k1 = huge_pinned(dm.N); hy, _k2 = huge_pinned(dm.M)
hx[:] = xh
for rnd in range(3):
    for name, bx, by in (("cudaHostAlloc", tx.numpy(), ty.numpy()), ("THP+register", hx, hy)):
        for _ in range(5): sp.spmv_host(sp.ELL_ROWS, dm, bx, by)
        t = time.perf_counter()
        for _ in range(150): sp.spmv_host(sp.ELL_ROWS, dm, bx, by)
        print("%-14s %.3f ms per call" % (name, (time.perf_counter() - t) / 150 * 1e3), flush=True)
print("equal:", np.array_equal(ty.numpy(), hy))
