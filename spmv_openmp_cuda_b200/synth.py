"""Seeded synthetic workloads of BASELINE.json, host and device builders.

Host builders return an engine.Spmat in the reference's layout (64-bit ids) -- what the CPU oracle
and the reference's OpenMP kernels consume; device builders return an engine.DeviceSpmat built
directly on the GPU.  Both call the same counter-based generator (csrc/synth.cu), so the matrices
are identical and any row range can be produced on its own (multi-GPU slices, bounded CPU samples).
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import Synth, check, lib, ptr
from .engine import DeviceSpmat, Spmat

SEED_BASE = 0x5EED0000
X_SCALE = 3e-5  # |x_i| < 3e-5 like the reference's MAXRND (src/include/config.h:115), but finite and seeded


def lap2d(n, seed=SEED_BASE + 1):
    """cfg1: 5-point 2-D Laplacian on an n x n grid (diag 4, off-diag -1)."""
    return Synth(1, seed, n, 0, 0, 0)


def stencil27(nx, ny=None, nz=None, seed=SEED_BASE + 2):
    """cfg2: 27-point 3-D stencil (diag 26, others -1)."""
    return Synth(2, seed, nx, ny or nx, nz or nx, 0)


def banded(M, nnz_per_row=32, half_width=1 << 15, seed=SEED_BASE + 4):
    """cfg4: random banded, nnz_per_row distinct sorted columns in [i-w, i+w] (stratified draw)."""
    return Synth(4, seed, half_width, nnz_per_row, M, 0)


def mixed(M, k_max, p_long, seed=SEED_BASE + 5):
    """cfg5: rows of length 4, a fraction p_long of rows of length k_max, uniform random columns."""
    return Synth(5, seed, M, k_max, 0, int(round(p_long * (1 << 32))))


def dims(s):
    M, N = C.c_uint64(), C.c_uint64()
    check(lib().spmvb200_synth_dims(C.byref(s), C.byref(M), C.byref(N)), "synth_dims")
    return M.value, N.value


def host_csr(s, row_begin=0, row_end=None):
    """Rows [row_begin,row_end) as a host CSR Spmat (local row numbering, global column ids)."""
    M, N = dims(s)
    row_end = M if row_end is None else row_end
    rows = row_end - row_begin
    rl = np.zeros(max(rows, 1), dtype=np.uint64)[:rows]
    check(lib().spmvb200_synth_rowlen_host(C.byref(s), row_begin, row_end, ptr(rl) if rows else None), "synth_rowlen_host") if rows else None
    irp = np.zeros(rows + 1, dtype=np.uint64)
    np.cumsum(rl, out=irp[1:])
    nz = int(irp[-1])
    ja = np.zeros(max(nz, 1), dtype=np.uint64)
    as_ = np.zeros(max(nz, 1), dtype=np.float64)
    if rows:
        check(lib().spmvb200_synth_fill_host(C.byref(s), row_begin, row_end, ptr(irp), ptr(ja), ptr(as_)), "synth_fill_host")
    return Spmat(rows, N, nz, ja[:nz], as_[:nz], IRP=irp, RL=rl, MAX_ROW_NZ=int(rl.max()) if rows else 0)


def device_csr(s, row_begin=0, row_end=None):
    capi.require_device()
    M, _ = dims(s)
    row_end = M if row_end is None else row_end
    out = C.c_void_p()
    check(lib().spmvb200_synth_csr_device(C.byref(s), row_begin, row_end, C.byref(out)), "synth_csr_device")
    return DeviceSpmat(out.value)


def rmat_host_csr(scale, edge_factor=16, seed=SEED_BASE + 3):
    """cfg3: R-MAT (a,b,c,d = .57,.19,.19,.05), duplicates merged, values hashed from (row, col)."""
    n_edges = edge_factor << scale
    keys = np.empty(n_edges, dtype=np.uint64)
    check(lib().spmvb200_synth_rmat_keys_host(scale, seed, 0, n_edges, ptr(keys)), "rmat_keys_host")
    keys = np.unique(keys)
    M = 1 << scale
    rows = (keys >> np.uint64(32)).astype(np.int64)
    ja = (keys & np.uint64(0xffffffff)).astype(np.uint64)
    as_ = np.empty(len(keys), dtype=np.float64)
    check(lib().spmvb200_synth_rmat_values_host(seed, len(keys), ptr(keys), ptr(as_)), "rmat_values_host")
    irp = np.zeros(M + 1, dtype=np.uint64)
    np.cumsum(np.bincount(rows, minlength=M), out=irp[1:])
    return Spmat.csr(M, irp, ja, as_)


def rmat_device_csr(scale, edge_factor=16, seed=SEED_BASE + 3):
    capi.require_device()
    out = C.c_void_p()
    check(lib().spmvb200_synth_rmat_csr_device(scale, edge_factor << scale, seed, C.byref(out)), "rmat_csr_device")
    return DeviceSpmat(out.value)


def host_vector(n, seed=SEED_BASE + 0x77, begin=0, scale=X_SCALE):
    x = np.empty(n, dtype=np.float64)
    if n:
        check(lib().spmvb200_synth_vector_host(seed, begin, begin + n, scale, ptr(x)), "synth_vector_host")
    return x


def device_vector_fill(d_x, n, seed=SEED_BASE + 0x77, begin=0, scale=X_SCALE):
    check(lib().spmvb200_synth_vector_device(seed, begin, begin + n, scale, ptr(d_x)), "synth_vector_device")


def csr_to_ell_host(mat):
    """Host CSR Spmat -> host row-major ELL Spmat with the reference's padding (AS=0, JA=0,
    src/lib/parser.c:245-252).  numpy only -- a data-format helper for tests, not a compute path."""
    rl = np.diff(mat.IRP).astype(np.int64)
    K = int(rl.max()) if mat.M else 0
    if mat.M and int(rl.min()) == K:  # every row full (cfg4: 32 per row): the row-major ELL arrays ARE the CSR arrays
        return Spmat.ell(mat.M, mat.N, K, mat.JA, mat.AS, RL=rl.astype(np.uint64), NZ=mat.NZ)
    ja = np.zeros(mat.M * K, dtype=np.uint64)
    as_ = np.zeros(mat.M * K, dtype=np.float64)
    if mat.NZ:
        row_of = np.repeat(np.arange(mat.M, dtype=np.int64), rl)
        slot = np.arange(mat.NZ, dtype=np.int64) - np.repeat(mat.IRP[:-1].astype(np.int64), rl)
        ja[row_of * K + slot] = mat.JA
        as_[row_of * K + slot] = mat.AS
    return Spmat.ell(mat.M, mat.N, K, ja, as_, RL=rl.astype(np.uint64), NZ=mat.NZ)


# the five BASELINE.json configurations at full size, and small versions for parity tests
FULL = {
    "cfg1_lap2d_1024": lambda: lap2d(1024),
    "cfg2_stencil27_128": lambda: stencil27(128),
    "cfg4_banded_2p25": lambda: banded(1 << 25, 32, 1 << 15),
}
