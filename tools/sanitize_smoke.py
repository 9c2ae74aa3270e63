#!/usr/bin/env python
"""Small end-to-end run of every kernel kind for compute-sanitizer (memcheck / racecheck): tiny matrices
including rows longer than a tile, empty rows, the pipelined host path and the adaptive tuner."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spmv_openmp_cuda_b200 as sp
from spmv_openmp_cuda_b200 import synth
import oracle

def check(mat, tag):
    x = synth.host_vector(mat.N)
    yr = oracle.sgemv_serial(mat.IRP, mat.JA, mat.AS, x)
    ell = synth.csr_to_ell_host(mat) if mat.M * mat.MAX_ROW_NZ < 4e6 else None
    dx, dy = sp.DeviceVector.from_host(x), sp.DeviceVector(mat.M)
    d = sp.spMatCpyCSR(mat)
    hs = [(f, d) for f in sp.SpmvCUDA_CSRFuncs]
    if ell is not None:
        hs += [(sp.cudaSpMVRowsELL, sp.spMatCpyELL(ell)), (sp.cudaSpMVRowsELLNNTransposed, sp.spMatCpyELLNNPitched(ell)),
               (sp.cudaSpMVWarpsPerRowELLNTrasposed, sp.spMatCpyELLNNPitched(ell))]
    for f, h in hs:
        dy.fill_bytes(0xFF); f(h, dx, sp.Config(), dy)
        bad, worst = oracle.strict_diff_csr(mat.IRP, mat.JA, mat.AS, x, yr, dy.to_host(), 1e-12)
        assert bad == 0, (tag, f.__name__, worst)
    for f in sp.SpmvB200CSRFuncs:
        y = np.full(mat.M, np.nan); f(mat, x, sp.Config(), y)
        assert oracle.strict_diff_csr(mat.IRP, mat.JA, mat.AS, x, yr, y, 1e-12)[0] == 0, (tag, "host")
    sp.cache_drop()
    print("ok", tag, mat.M, mat.NZ)

check(synth.host_csr(synth.lap2d(40)), "lap2d")
check(synth.rmat_host_csr(11, 16), "rmat")
check(synth.host_csr(synth.banded(70000, 32, 900)), "banded(pipelined host path)")
rng = np.random.default_rng(1)
lens = np.r_[np.zeros(50, int), 5000, np.full(300, 3), 2300, 0, 0]
irp = np.zeros(len(lens) + 1, dtype=np.uint64); irp[1:] = np.cumsum(lens)
ja = np.concatenate([np.sort(rng.choice(6000, k, replace=False)) for k in lens]).astype(np.uint64)
check(sp.Spmat.csr(6000, irp, ja, rng.uniform(-1, 1, int(irp[-1]))), "long rows + empty rows")
print("SANITIZE_SMOKE_OK")
