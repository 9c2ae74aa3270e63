"""Exact CSR stream kernel (SPMVB200_CSR_ROWS, re-tiled copies off): time per SpMV against the product pre-pass threshold.
    for t in 0 8 16 32 48 64 128 100000; do SPMVB200_PRE_T=$t python tools/pre_t_sweep.py; done"""
import os, sys
os.environ["SPMVB200_EXACT_ONLY_STREAM"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spmv_openmp_cuda_b200 as sp
from spmv_openmp_cuda_b200 import synth

out = []
for name, make in (("cfg2", lambda: synth.device_csr(synth.stencil27(128))), ("cfg3", lambda: synth.rmat_device_csr(22, 16)),
                   ("cfg5 K32 p.15", lambda: synth.device_csr(synth.mixed(1 << 23, 32, 0.15))), ("cfg5 K48 p1", lambda: synth.device_csr(synth.mixed(1 << 22, 48, 1.0))),
                   ("cfg4s", lambda: synth.device_csr(synth.banded(1 << 22, 32, 1 << 15)))):
    if len(sys.argv) > 1 and not any(name.startswith(a) for a in sys.argv[1:]):
        continue
    d = make()
    dx = sp.DeviceVector(d.N); synth.device_vector_fill(dx, d.N); dy = sp.DeviceVector(d.M)
    sp.time_kernel(sp.CSR_ROWS, d, dx, dy, reps=3)
    t = sp.time_kernel(sp.CSR_ROWS, d, dx, dy, reps=15)
    out.append("%s %.1f" % (name, float(t.mean()) * 1e3))
    d.free()
print("pre_t=%-7s" % os.environ.get("SPMVB200_PRE_T", "default"), " | ".join(out), "(us)", flush=True)
