"""ctypes bindings for the oracle (liboracle.so) and the compiled reference (_ref).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Index arrays are numpy uint64 (the
reference's `ulong`, src/include/sparseMatrix.h:26-32), values float64.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libspmv_ref.so")

_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(ref=True):
    """Compile the restatement (always) and the reference (only where /root/reference exists)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = C.CDLL(ORACLE_SO)
        L.oracle_sgemv_serial.argtypes = [C.c_uint64, _u64p, _u64p, _f64p, _f64p, _f64p]
        L.oracle_spmv_rows_blocks_csr.argtypes = [C.c_uint64, _u64p, _u64p, _f64p, _f64p, _f64p,
                                                  C.c_uint, C.c_int]
        L.oracle_spmv_rows_basic_csr.argtypes = [C.c_uint64, _u64p, _u64p, _f64p, _f64p, _f64p, C.c_int]
        L.oracle_spmv_rows_ell.argtypes = [C.c_uint64, C.c_uint64, _u64p, _f64p, C.c_void_p, _f64p, _f64p]
        L.oracle_spmv_ell_colmajor.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, _u64p, _f64p, _f64p, _f64p]
        L.oracle_coo_to_csr.argtypes = [C.c_uint64, C.c_uint64, _u64p, _u64p, _f64p, _u64p, _u64p, _f64p,
                                        C.c_void_p, C.c_int]
        L.oracle_max_row_len.argtypes = [C.c_uint64, C.c_uint64, _u64p]
        L.oracle_max_row_len.restype = C.c_uint64
        L.oracle_coo_to_ell.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, _u64p, _u64p, _f64p, _u64p, _f64p,
                                        C.c_void_p]
        L.oracle_ell_transpose.argtypes = [C.c_uint64, C.c_uint64, _u64p, _f64p, _u64p, _f64p]
        L.oracle_double_vectors_diff.argtypes = [_f64p, _f64p, C.c_uint64, C.POINTER(C.c_double)]
        L.oracle_strict_diff_csr.argtypes = [C.c_uint64, _u64p, _u64p, _f64p, _f64p, _f64p, _f64p, C.c_double,
                                             C.POINTER(C.c_double)]
        L.oracle_strict_diff_csr.restype = C.c_uint64
        L.oracle_stats_avg_var.argtypes = [_f64p, C.c_uint, _f64p]
        L.oracle_omp_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# ---------------------------------------------------------------- restatement wrappers
def sgemv_serial(irp, ja, as_, x):
    irp, ja, as_, x = _c(irp, np.uint64), _c(ja, np.uint64), _c(as_, np.float64), _c(x, np.float64)
    y = np.empty(len(irp) - 1, dtype=np.float64)
    lib().oracle_sgemv_serial(len(y), irp, ja, as_, x, y)
    return y


def spmv_rows_blocks_csr(irp, ja, as_, x, grid_rows=8, threads=0, out=None):
    y = np.empty(len(irp) - 1, dtype=np.float64) if out is None else out
    rc = lib().oracle_spmv_rows_blocks_csr(len(y), irp, ja, as_, x, y, grid_rows, threads)
    if rc:
        raise RuntimeError("oracle_spmv_rows_blocks_csr failed")
    return y


def spmv_rows_basic_csr(irp, ja, as_, x, threads=0, out=None):
    y = np.empty(len(irp) - 1, dtype=np.float64) if out is None else out
    lib().oracle_spmv_rows_basic_csr(len(y), irp, ja, as_, x, y, threads)
    return y


def spmv_rows_ell(M, K, ja, as_, x, rl=None):
    y = np.empty(M, dtype=np.float64)
    rlp = None if rl is None else _c(rl, np.uint64).ctypes.data_as(C.c_void_p)
    if rl is not None:
        rl = _c(rl, np.uint64)
        rlp = rl.ctypes.data_as(C.c_void_p)
    lib().oracle_spmv_rows_ell(M, K, _c(ja, np.uint64), _c(as_, np.float64), rlp, _c(x, np.float64), y)
    return y


def spmv_ell_colmajor(M, K, pitch, ja_t, as_t, x):
    y = np.empty(M, dtype=np.float64)
    lib().oracle_spmv_ell_colmajor(M, K, pitch, _c(ja_t, np.uint64), _c(as_t, np.float64), _c(x, np.float64), y)
    return y


def coo_to_csr(M, row, col, val, check_sorted=True):
    row, col, val = _c(row, np.uint64), _c(col, np.uint64), _c(val, np.float64)
    nz = len(row)
    irp = np.zeros(M + 1, dtype=np.uint64)
    ja = np.zeros(max(nz, 1), dtype=np.uint64)
    as_ = np.zeros(max(nz, 1), dtype=np.float64)
    rl = np.zeros(max(M, 1), dtype=np.uint64)
    rc = lib().oracle_coo_to_csr(nz, M, row, col, val, irp, ja, as_, rl.ctypes.data_as(C.c_void_p),
                                 int(check_sorted))
    if rc:
        raise ValueError("COO entries not column-sorted within rows" if rc == 1 else "alloc failure")
    return irp, ja[:nz], as_[:nz], rl[:M]


def coo_to_ell(M, row, col, val):
    row, col, val = _c(row, np.uint64), _c(col, np.uint64), _c(val, np.float64)
    nz = len(row)
    K = int(lib().oracle_max_row_len(nz, M, row))
    ja = np.zeros(max(M * K, 1), dtype=np.uint64)
    as_ = np.zeros(max(M * K, 1), dtype=np.float64)
    rl = np.zeros(max(M, 1), dtype=np.uint64)
    rc = lib().oracle_coo_to_ell(nz, M, K, row, col, val, ja, as_, rl.ctypes.data_as(C.c_void_p))
    if rc:
        raise ValueError("ELL size cap exceeded" if rc == 3 else "alloc failure")
    return K, ja[:M * K], as_[:M * K], rl[:M]


def ell_transpose(M, K, ja, as_):
    ja_t = np.empty(M * K, dtype=np.uint64)
    as_t = np.empty(M * K, dtype=np.float64)
    lib().oracle_ell_transpose(M, K, _c(ja, np.uint64), _c(as_, np.float64), ja_t, as_t)
    return ja_t, as_t


def double_vectors_diff(a, b):
    """Reference comparator: (failed?, signed max diff).  abs threshold 7e-4, NaN-blind."""
    dm = C.c_double(0)
    rc = lib().oracle_double_vectors_diff(_c(a, np.float64), _c(b, np.float64), len(a), C.byref(dm))
    return bool(rc), dm.value


def strict_diff_csr(irp, ja, as_, x, yref, y, tau=1e-12):
    """(number of failing rows, worst |dy| / sum|a||x|).  NaN-aware."""
    w = C.c_double(0)
    bad = lib().oracle_strict_diff_csr(len(irp) - 1, _c(irp, np.uint64), _c(ja, np.uint64), _c(as_, np.float64),
                                       _c(x, np.float64), _c(yref, np.float64), _c(y, np.float64), tau, C.byref(w))
    return int(bad), w.value


def stats_avg_var(v):
    out = np.zeros(2)
    v = _c(v, np.float64)
    lib().oracle_stats_avg_var(v, len(v), out)
    return out[0], out[1]


def omp_max_threads():
    return int(lib().oracle_omp_max_threads())


# ---------------------------------------------------------------- the compiled reference
class RefSpmat(C.Structure):
    """`spmat` as gcc sees it with -DROWLENS (src/include/sparseMatrix.h:25-42; no pitch fields
    because __CUDACC__ is undefined in C translation units -- SURVEY.md §2.3-10)."""
    _fields_ = [("NZ", C.c_ulong), ("M", C.c_ulong), ("N", C.c_ulong), ("JA", C.c_void_p), ("RL", C.c_void_p),
                ("IRP", C.c_void_p), ("MAX_ROW_NZ", C.c_ulong), ("AS", C.c_void_p)]


class RefConfig(C.Structure):
    """`CONFIG` as gcc sees it (src/include/config.h:21-32)."""
    _fields_ = [("gridRows", C.c_ushort), ("gridCols", C.c_ushort), ("threadNum", C.c_uint),
                ("chunkDistrbFunc", C.c_void_p)]


_ref = {}
# build variants of the same unmodified sources: "default" = the reference's own configuration (SIMD_ROWS_REDUCTION TRUE,
# src/include/config.h:92-94), "nosimd" = -DSIMD_ROWS_REDUCTION=FALSE (the other setting SURVEY.md §8d asks the CPU baseline to try)
REF_VARIANTS = {"default": REF_SO, "nosimd": os.path.join(HERE, "_ref", "libspmv_ref_nosimd.so")}


def ref_available(variant="default"):
    return os.path.exists(REF_VARIANTS[variant])


def ref(variant="default"):
    """The unmodified reference, compiled by `make -C oracle ref`.  None-safe: raises if absent."""
    if variant not in _ref:
        if not ref_available(variant):
            raise FileNotFoundError(REF_VARIANTS[variant] + " (build with `make -C oracle ref` where /root/reference exists)")
        L = C.CDLL(REF_VARIANTS[variant])  # RTLD_LOCAL: the variants define the same symbols and must not interpose each other
        assert L.refshim_sizeof_spmat() == C.sizeof(RefSpmat), "spmat ABI mismatch"
        assert L.refshim_sizeof_config() == C.sizeof(RefConfig), "CONFIG ABI mismatch"
        assert L.refshim_rowlens() == 1
        L.refshim_chunks_fn.restype = C.c_void_p
        L.refshim_elapsed_internal.restype = C.c_double
        for name in ("sgemvSerial", "spmvRowsBasicCSR", "spmvRowsBlocksCSR", "spmvTilesCSR", "spmvTilesAllocdCSR",
                     "spmvRowsBasicELL", "spmvRowsBlocksELL", "spmvTilesELL"):
            f = getattr(L, name)
            f.argtypes = [C.POINTER(RefSpmat), _f64p, C.POINTER(RefConfig), _f64p]
            f.restype = C.c_int
        L.MMtoCSR.argtypes = [C.c_char_p]
        L.MMtoCSR.restype = C.POINTER(RefSpmat)
        L.MMtoELL.argtypes = [C.c_char_p]
        L.MMtoELL.restype = C.POINTER(RefSpmat)
        L.ellTranspose.argtypes = [C.POINTER(RefSpmat)]
        L.ellTranspose.restype = C.POINTER(RefSpmat)
        L.refshim_free_spmat.argtypes = [C.POINTER(RefSpmat)]
        L.doubleVectorsDiff.argtypes = [_f64p, _f64p, C.c_ulong, C.POINTER(C.c_double)]
        L.doubleVectorsDiff.restype = C.c_int
        _ref[variant] = L
    return _ref[variant]


def ref_spmat(M, N, nz, ja, as_, irp=None, rl=None, max_row_nz=0):
    """Wrap numpy arrays (kept alive by the caller) in a reference `spmat`."""
    m = RefSpmat()
    m.NZ, m.M, m.N, m.MAX_ROW_NZ = nz, M, N, max_row_nz
    m.JA = ja.ctypes.data
    m.AS = as_.ctypes.data
    m.IRP = irp.ctypes.data if irp is not None else None
    m.RL = rl.ctypes.data if rl is not None else None
    return m


def ref_config(grid_rows=8, grid_cols=8, threads=None, chunks=0, variant="default"):
    """chunks: 0 = chunksNOOP, 1 = chunksFair, 2 = chunksFairFolded (ompChunksDivide.h:33-91).  The chunk function pointer belongs
    to one build variant: pass the variant the config will be used with."""
    c = RefConfig()
    c.gridRows, c.gridCols = grid_rows, grid_cols
    c.threadNum = threads or omp_max_threads()
    c.chunkDistrbFunc = ref(variant).refshim_chunks_fn(chunks)
    return c


def ref_call(name, mat, x, cfg, M, variant="default"):
    y = np.empty(M, dtype=np.float64)
    rc = getattr(ref(variant), name)(C.byref(mat), _c(x, np.float64), C.byref(cfg), y)
    if rc:
        raise RuntimeError("reference %s returned %d" % (name, rc))
    return y


def _arr(ptr, n, dt):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dt)
    buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dt).copy()


def ref_mm_to_csr(path):
    """Parse a Matrix Market file with the reference's own reader (src/lib/parser.c:298-345)."""
    p = ref().MMtoCSR(path.encode())
    if not p:
        raise RuntimeError("reference MMtoCSR failed on " + path)
    m = p.contents
    out = dict(M=m.M, N=m.N, NZ=m.NZ, irp=_arr(m.IRP, m.M + 1, np.uint64), ja=_arr(m.JA, m.NZ, np.uint64),
               as_=_arr(m.AS, m.NZ, np.float64), rl=_arr(m.RL, m.M, np.uint64))
    ref().refshim_free_spmat(p)
    return out


def ref_mm_to_ell(path, transpose=False):
    """Reference MMtoELL (parser.c:347-376), optionally followed by ellTranspose (sparseUtils.c:145-185)."""
    p = ref().MMtoELL(path.encode())
    if not p:
        raise RuntimeError("reference MMtoELL failed on " + path)
    m = p.contents
    M, K = m.M, m.MAX_ROW_NZ
    out = dict(M=M, N=m.N, NZ=m.NZ, K=K, ja=_arr(m.JA, M * K, np.uint64), as_=_arr(m.AS, M * K, np.float64),
               rl=_arr(m.RL, M, np.uint64))
    if transpose:
        t = ref().ellTranspose(p)
        tm = t.contents
        out["ja_t"] = _arr(tm.JA, M * K, np.uint64)
        out["as_t"] = _arr(tm.AS, M * K, np.float64)
        out["t_M"], out["t_N"], out["t_MAX_ROW_NZ"] = tm.M, tm.N, tm.MAX_ROW_NZ
        tm.IRP = None
        ref().refshim_free_spmat(t)
    m.IRP = None  # MMtoELL leaves IRP NULL (calloc'd struct); keep free() well defined
    ref().refshim_free_spmat(p)
    return out


# ---------------------------------------------------------------- Matrix Market -> COO (restatement)
def mm_to_coo(text):
    """Coordinate Matrix Market text -> (M, N, row, col, val) in the reference's COO order.

    Follows MMtoCOO, src/lib/parser.c:42-105: entries are kept in FILE order, indices become
    0-based, `pattern` files get val = 1.0 (:60-62), and for `symmetric` files the mirrored
    entry (col,row) is appended right after each off-diagonal entry (:85-90).
    """
    lines = text.splitlines()
    banner = lines[0].lower().split()
    assert banner[0] == "%%matrixmarket" and banner[1] == "matrix" and banner[2] == "coordinate", banner
    pattern, symmetric = banner[3] == "pattern", banner[4] == "symmetric"
    body = [ln for ln in lines[1:] if ln.strip() and not ln.lstrip().startswith("%")]
    M, N, nz = (int(t) for t in body[0].split())
    row, col, val = [], [], []
    for ln in body[1:1 + nz]:
        t = ln.split()
        r, c = int(t[0]) - 1, int(t[1]) - 1
        v = 1.0 if pattern else float(t[2])
        row.append(r); col.append(c); val.append(v)
        if symmetric and r != c:
            row.append(c); col.append(r); val.append(v)
    return M, N, np.array(row, dtype=np.uint64), np.array(col, dtype=np.uint64), np.array(val, dtype=np.float64)
