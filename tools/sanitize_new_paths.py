#!/usr/bin/env python
"""Stand-alone check (plain, or under compute-sanitizer where the pool allows it -- this one does not) of the paths that need >= 2^20
non-zeros to be taken: the exact kind's SELL hybrid with the serial-order row kernel (R-MAT) and ELL_ROWS' SELL copy built from the ELL
arrays (mixed short / long rows).  The same cases are in tests/test_gpu_parity.py."""
import os, sys
os.environ["SPMVB200_FORCE_EXACT"] = "13"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spmv_openmp_cuda_b200 as sp
from spmv_openmp_cuda_b200 import synth
import oracle

mat = synth.rmat_host_csr(17, 16)
x = synth.host_vector(mat.N)
yr = oracle.sgemv_serial(mat.IRP, mat.JA, mat.AS, x)
dx, dy, d = sp.DeviceVector.from_host(x), sp.DeviceVector(mat.M), sp.spMatCpyCSR(mat)
dy.fill_bytes(0xFF); sp.cudaSpMVRowsCSR(d, dx, sp.Config(), dy)
lens = np.diff(mat.IRP)
y = dy.to_host()
assert d.exact_choice == "sell" and np.array_equal(y[lens <= 2048], yr[lens <= 2048])
print("ok exact hybrid", mat.M, mat.NZ, flush=True)
mat = synth.host_csr(synth.mixed(300000, 32, 0.05))
ell = synth.csr_to_ell_host(mat)
x = synth.host_vector(mat.N)
yr = oracle.sgemv_serial(mat.IRP, mat.JA, mat.AS, x)
dx, dy, e = sp.DeviceVector.from_host(x), sp.DeviceVector(mat.M), sp.spMatCpyELL(ell)
dy.fill_bytes(0xFF); sp.cudaSpMVRowsELL(e, dx, sp.Config(), dy)
assert np.array_equal(dy.to_host(), yr)
print("ok ell ->", e.exact_choice, flush=True)
print("SANITIZE_NEW_PATHS_OK")
