"""Host-side mirror of the reference's GPU interface, on top of the C ABI (include/spmv_b200.h).

Names, argument order and error behaviour follow the reference so that the parity tests read like
its harness (test/SpMV_test.cu):

  reference (C / CUDA)                                   here
  ---------------------------------------------------    ------------------------------------------
  spmat  (src/include/sparseMatrix.h:25-42)              Spmat           (host arrays, 64-bit ids)
  CONFIG (src/include/config.h:21-32)                    Config
  spMatCpyCSR / spMatCpyELL (src/commons/cudaUtils.cu)   spMatCpyCSR / spMatCpyELL -> DeviceSpmat
  cudaFreeSpmat (src/include/cudaUtils.h:70-78)          cudaFreeSpmat
  f<<<grid,block>>>(dMat,dVect,Conf,dOutV)               f(dMat, dVect, Conf, dOutV)  with f one of
  SpmvCUDA_CSRFuncs / SpmvCUDA_ELLFuncs (SpMV.h:130-142) the same-named functions / tables below
  SPMV_INTERF f(mat,x,&Conf,y) (SpMV.h:63-64)            b200SpMV*(mat, x, Conf, y)  (host buffers)

Return convention: like the reference, functions return 0 (EXIT_SUCCESS); failures raise
SpmvB200Error carrying the library's diagnostic (the C layer itself returns non-zero + stderr).
There is no CPU fallback anywhere in this package.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import (CSR_ADAPTIVE, CSR_ROWS, CSR_ROWS_WARP, ELL_ROWS, ELL_ROWS_NT, ELL_ROWS_WARP_NT, FMT_CSR,
                   FMT_ELL_COLMAJOR, FMT_ELL_ROWMAJOR, FMT_SELL, FMT_XWIN, SELL_ROWS, XWIN_ROWS, SpmvB200Error, check, lib, ptr)

EXIT_SUCCESS = 0

# compute-mode strings, src/include/SpMV.h:37-41 (+ the appended adaptive mode)
CUDA_CSR_ROWS = "CUDA_CSR_ROWS"
CUDA_CSR_ROWS_WARP = "CUDA_CSR_ROWS_WARP"
CUDA_ELL_ROWS = "CUDA_ELL_ROWS"
CUDA_ELL_ROWS_WARP = "CUDA_ELL_ROWS_WARP"
CUDA_ELL_ROWS_WARP_NT = "CUDA_ELL_ROWS_WARP_NN_TRANSPOSED"
CUDA_CSR_ADAPTIVE = "CUDA_CSR_ADAPTIVE"


class Spmat:
    """Host sparse matrix with the reference's field names (src/include/sparseMatrix.h:25-42).
    CSR: IRP[M+1], JA[NZ], AS[NZ] (+RL[M]).  ELL: JA/AS row-major M x MAX_ROW_NZ (+RL[M]), IRP None."""

    def __init__(self, M, N, NZ, JA, AS, IRP=None, RL=None, MAX_ROW_NZ=0):
        self.M, self.N, self.NZ, self.MAX_ROW_NZ = int(M), int(N), int(NZ), int(MAX_ROW_NZ)
        self.JA = np.ascontiguousarray(JA, dtype=np.uint64)
        self.AS = np.ascontiguousarray(AS, dtype=np.float64)
        self.IRP = None if IRP is None else np.ascontiguousarray(IRP, dtype=np.uint64)
        self.RL = None if RL is None else np.ascontiguousarray(RL, dtype=np.uint64)

    @classmethod
    def csr(cls, N, IRP, JA, AS, RL=None):
        IRP = np.ascontiguousarray(IRP, dtype=np.uint64)
        M = len(IRP) - 1
        if RL is None:
            RL = np.diff(IRP)
        return cls(M, N, int(IRP[-1]), JA, AS, IRP=IRP, RL=RL, MAX_ROW_NZ=int(np.max(np.diff(IRP))) if M else 0)

    @classmethod
    def ell(cls, M, N, K, JA, AS, RL=None, NZ=None):
        if NZ is None:
            NZ = int(np.sum(RL)) if RL is not None else int(np.count_nonzero(AS))
        return cls(M, N, NZ, JA, AS, IRP=None, RL=RL, MAX_ROW_NZ=K)


    # ---- binary cache next to the Matrix Market text: parsing a 1e8-entry .mtx takes minutes, this takes a read()
    _MAGIC = b"SPMVB2\x00\x01"

    def save(self, path):
        """Raw little-endian dump: magic, (M, N, NZ, MAX_ROW_NZ, has_irp, has_rl) as uint64, then IRP / JA / AS / RL."""
        with open(path, "wb") as f:
            f.write(self._MAGIC)
            np.array([self.M, self.N, self.NZ, self.MAX_ROW_NZ, self.IRP is not None, self.RL is not None], dtype="<u8").tofile(f)
            for a in (self.IRP, self.JA, self.AS, self.RL):
                if a is not None:
                    np.array([a.size], dtype="<u8").tofile(f)
                    a.tofile(f)

    @classmethod
    def load(cls, path):
        with open(path, "rb") as f:
            if f.read(8) != cls._MAGIC:
                raise ValueError("%s is not a spmv_b200 binary matrix" % path)
            M, N, NZ, K, has_irp, has_rl = (int(v) for v in np.fromfile(f, dtype="<u8", count=6))

            def arr(dtype):
                n = int(np.fromfile(f, dtype="<u8", count=1)[0])
                a = np.fromfile(f, dtype=dtype, count=n)
                if a.size != n:
                    raise ValueError("%s is truncated" % path)
                return a
            IRP = arr("<u8") if has_irp else None
            JA, AS = arr("<u8"), arr("<f8")
            RL = arr("<u8") if has_rl else None
        return cls(M, N, NZ, JA, AS, IRP=IRP, RL=RL, MAX_ROW_NZ=K)


class Config:
    """CONFIG, src/include/config.h:21-32.  gridSize/blockSize are accepted and ignored by the engine
    (it chooses its own launch geometry)."""

    def __init__(self, gridRows=8, gridCols=8, threadNum=0, gridSize=None, blockSize=None, sharedMemSize=0):
        self.gridRows, self.gridCols, self.threadNum = gridRows, gridCols, threadNum
        self.gridSize, self.blockSize, self.sharedMemSize = gridSize, blockSize, sharedMemSize


class DeviceSpmat:
    """Device-resident matrix: owns one opaque spmvb200_matrix handle."""

    def __init__(self, handle):
        self._h = handle
        M, N, NZ, K, fmt = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_int()
        check(lib().spmvb200_dims(handle, C.byref(M), C.byref(N), C.byref(NZ), C.byref(K), C.byref(fmt)), "dims")
        self.M, self.N, self.NZ, self.MAX_ROW_NZ, self.format = M.value, N.value, NZ.value, K.value, fmt.value

    @property
    def handle(self):
        if not self._h:
            raise SpmvB200Error("matrix handle already freed")
        return self._h

    @property
    def algorithmic_bytes(self):
        return int(lib().spmvb200_algorithmic_bytes(self.handle))

    @property
    def device_bytes(self):
        return int(lib().spmvb200_device_bytes(self.handle))

    @property
    def adaptive_choice(self):
        buf = C.create_string_buffer(64)
        check(lib().spmvb200_adaptive_choice(self.handle, buf, 64), "adaptive_choice")
        return buf.value.decode()

    @property
    def exact_choice(self):
        """what SPMVB200_CSR_ROWS / SPMVB200_ELL_ROWS picked at first use: stream | ell | xwindow | sell ("" before)"""
        buf = C.create_string_buffer(64)
        check(lib().spmvb200_exact_choice(self.handle, buf, 64), "exact_choice")
        return buf.value.decode()

    @property
    def index_bits(self):
        return int(lib().spmvb200_index_bits(self.handle))

    def supports(self, kind):
        return bool(lib().spmvb200_kind_supported(self.handle, kind))

    def to_ell(self, fmt=FMT_ELL_COLMAJOR):
        out = C.c_void_p()
        check(lib().spmvb200_ell_from_csr(self.handle, fmt, C.byref(out)), "ell_from_csr")
        return DeviceSpmat(out.value)

    def to_sell(self, sigma=0):
        """SELL-32-sigma copy of a CSR handle (rows sorted by length inside windows of sigma rows)."""
        out = C.c_void_p()
        check(lib().spmvb200_sell_from_csr(self.handle, sigma, C.byref(out)), "sell_from_csr")
        return DeviceSpmat(out.value)

    def to_xwin(self, rows_per_block=0, window_cols=0):
        """x-window copy of a CSR handle: row blocks x column windows, x gathered from shared memory (csrc/xwin.cuh)."""
        out = C.c_void_p()
        check(lib().spmvb200_xwin_from_csr(self.handle, rows_per_block, window_cols, C.byref(out)), "xwin_from_csr")
        return DeviceSpmat(out.value)

    @property
    def xwin_info(self):
        R, W, nt, ring, moved = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint64()
        check(lib().spmvb200_xwin_info(self.handle, C.byref(R), C.byref(W), C.byref(nt), C.byref(ring), C.byref(moved)), "xwin_info")
        return dict(rows_per_block=R.value, window_cols=W.value, ntiles=nt.value, ring=ring.value, moved_bytes=moved.value)

    @property
    def col_range(self):
        """(smallest, largest) column id referenced -- which part of x these rows read."""
        lo, hi = C.c_uint64(), C.c_uint64()
        check(lib().spmvb200_col_range(self.handle, C.byref(lo), C.byref(hi)), "col_range")
        return lo.value, hi.value

    def download_csr(self):
        irp = np.empty(self.M + 1, dtype=np.uint64)
        ja = np.empty(self.NZ, dtype=np.uint64)
        as_ = np.empty(self.NZ, dtype=np.float64)
        check(lib().spmvb200_csr_download(self.handle, ptr(irp), ptr(ja), ptr(as_)), "csr_download")
        return irp, ja, as_

    def free(self):
        if self._h:
            h, self._h = self._h, None
            check(lib().spmvb200_free(h), "free")

    def __del__(self):
        try:
            self.free()
        except Exception:  # noqa: BLE001
            pass


class DeviceVector:
    """fp64 device vector allocated through the C ABI (cudaMalloc)."""

    def __init__(self, n):
        self.n = int(n)
        p = C.c_void_p()
        check(lib().spmvb200_dmalloc(C.byref(p), max(self.n, 1) * 8), "dmalloc")
        self._p = p.value

    @classmethod
    def from_host(cls, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        v = cls(len(a))
        if len(a):
            check(lib().spmvb200_h2d(v._p, ptr(a), a.nbytes), "h2d")
        return v

    def data_ptr(self):
        return self._p

    def to_host(self, out=None):
        out = np.empty(self.n, dtype=np.float64) if out is None else out
        if self.n:
            check(lib().spmvb200_d2h(ptr(out), self._p, self.n * 8), "d2h")
        return out

    def fill_bytes(self, byte):
        a = np.frombuffer(bytes([byte]) * (self.n * 8), dtype=np.float64)
        if self.n:
            check(lib().spmvb200_h2d(self._p, ptr(a), a.nbytes), "h2d")

    def free(self):
        if self._p:
            p, self._p = self._p, None
            lib().spmvb200_dfree(p)

    def __del__(self):
        try:
            self.free()
        except Exception:  # noqa: BLE001
            pass


# ---------------------------------------------------------------------------------------- upload
def spMatCpyCSR(mat, row_begin=0, row_end=None):
    """Host CSR -> device (replaces spMatCpyCSR, src/commons/cudaUtils.cu:20-55).  A row range uploads
    one GPU's slice of a row-block partition."""
    capi.require_device()
    row_end = mat.M if row_end is None else row_end
    out = C.c_void_p()
    check(lib().spmvb200_csr_upload(mat.M, mat.N, ptr(mat.IRP), ptr(mat.JA), ptr(mat.AS), row_begin, row_end,
                                    C.byref(out)), "spMatCpyCSR")
    return DeviceSpmat(out.value)


def spMatCpyELL(mat, row_begin=0, row_end=None, use_rowlens=True):
    """Row-major host ELL -> pitched COLUMN-major device ELL; the transposition the reference does on
    the host (ellTranspose, src/commons/sparseUtils.c:145-185) happens on the device
    (replaces spMatCpyELL, src/commons/cudaUtils.cu:56-98)."""
    return _ell_upload(mat, FMT_ELL_COLMAJOR, row_begin, row_end, use_rowlens)


def spMatCpyELLNNPitched(mat, row_begin=0, row_end=None, use_rowlens=True):
    """Row-major host ELL -> row-major device ELL (replaces spMatCpyELLNNPitched,
    src/commons/cudaUtils.cu:100-140, and spMatCpyELL on the non-transposed matrix)."""
    return _ell_upload(mat, FMT_ELL_ROWMAJOR, row_begin, row_end, use_rowlens)


def _ell_upload(mat, fmt, row_begin, row_end, use_rowlens):
    capi.require_device()
    row_end = mat.M if row_end is None else row_end
    out = C.c_void_p()
    rl = mat.RL if use_rowlens else None
    check(lib().spmvb200_ell_upload(mat.M, mat.N, mat.MAX_ROW_NZ, ptr(mat.JA), ptr(mat.AS), ptr(rl), row_begin,
                                    row_end, fmt, C.byref(out)), "spMatCpyELL")
    return DeviceSpmat(out.value)


def cudaFreeSpmat(dmat):
    dmat.free()
    return EXIT_SUCCESS


# ---------------------------------------------------------------------------------------- kernels
def _launch(kind, m, v, cfg, outV, stream=None):
    check(lib().spmvb200_spmv_device(m.handle, kind, ptr(v), ptr(outV), stream), capi.lib().spmvb200_kind_name(kind).decode())
    return EXIT_SUCCESS


def cudaSpMVRowsCSR(m, v, cfg, outV, stream=None):
    """src/SpMV_CUDA.cu:33-49 -> TMA-staged CSR stream kernel / x-window or SELL copy (picked at first use), bit-identical to
    sgemvSerial for every row of at most 2048 non-zeros; a longer row is split into 2048-entry segments combined in segment order
    (deterministic, within 1e-12 * sum|a_ij x_j| of the serial sum, not the same bits)."""
    return _launch(CSR_ROWS, m, v, cfg, outV, stream)


def cudaSpMVWarpPerRowCSR(m, v, cfg, outV, stream=None):
    """src/SpMV_CUDA.cu:52-73 -> sub-warp-per-row vector kernel (128-bit loads, shuffle reduction)."""
    return _launch(CSR_ROWS_WARP, m, v, cfg, outV, stream)


def cudaSpMVAdaptiveCSR(m, v, cfg, outV, stream=None):
    """new mode: row-length adaptive stream kernel (long rows reduced by warps / split across CTAs)."""
    return _launch(CSR_ADAPTIVE, m, v, cfg, outV, stream)


def cudaSpMVRowsELL(m, v, cfg, outV, stream=None):
    """src/SpMV_CUDA.cu:79-96 -> column-major pitched ELL, row-length early exit."""
    return _launch(ELL_ROWS, m, v, cfg, outV, stream)


def cudaSpMVRowsSELL(m, v, cfg, outV, stream=None):
    """new mode: sliced ELL (SELL-32-sigma), thread per row, bit-identical to sgemvSerial for rows of at most 2048 non-zeros (rows
    longer than 256 run on the serial-order per-row kernels next to the slices; beyond 2048 they are split into segments)."""
    return _launch(SELL_ROWS, m, v, cfg, outV, stream)


def cudaSpMVRowsXWIN(m, v, cfg, outV, stream=None):
    """new mode: x-window CSR (x gathered from shared-memory windows), bit-identical to sgemvSerial for sorted rows."""
    return _launch(XWIN_ROWS, m, v, cfg, outV, stream)


def cudaSpMVRowsELLNNTransposed(m, v, cfg, outV, stream=None):
    """src/SpMV_CUDA.cu:99-115 -> row-major ELL, sub-warp per row sized from K."""
    return _launch(ELL_ROWS_NT, m, v, cfg, outV, stream)


def cudaSpMVWarpsPerRowELLNTrasposed(m, v, cfg, outV, stream=None):
    """src/SpMV_CUDA.cu:116-135 -> row-major ELL, one warp per row."""
    return _launch(ELL_ROWS_WARP_NT, m, v, cfg, outV, stream)


# function tables, src/include/SpMV.h:130-142 (the adaptive mode appended)
SpmvCUDA_CSRFuncs = [cudaSpMVRowsCSR, cudaSpMVWarpPerRowCSR, cudaSpMVAdaptiveCSR]  # SELL needs its own handle: see to_sell()
SpmvCUDA_CSRFuncs_WarpPerRowIdx = 1
SpmvCUDA_ELLFuncs = [cudaSpMVRowsELL, cudaSpMVRowsELLNNTransposed, cudaSpMVWarpsPerRowELLNTrasposed]
SpmvCUDA_ELLFuncs_NN_TraposedImpl = 1
SpmvCUDA_ELLFuncs_WarpPerRowIdx = 2

KIND_OF = {cudaSpMVRowsCSR: CSR_ROWS, cudaSpMVWarpPerRowCSR: CSR_ROWS_WARP, cudaSpMVAdaptiveCSR: CSR_ADAPTIVE,
           cudaSpMVRowsSELL: SELL_ROWS, cudaSpMVRowsXWIN: XWIN_ROWS, cudaSpMVRowsELL: ELL_ROWS, cudaSpMVRowsELLNNTransposed: ELL_ROWS_NT,
           cudaSpMVWarpsPerRowELLNTrasposed: ELL_ROWS_WARP_NT}
MODE_OF = {CUDA_CSR_ROWS: CSR_ROWS, CUDA_CSR_ROWS_WARP: CSR_ROWS_WARP, CUDA_ELL_ROWS: ELL_ROWS,
           CUDA_ELL_ROWS_WARP_NT: ELL_ROWS_WARP_NT, CUDA_CSR_ADAPTIVE: CSR_ADAPTIVE}


def time_kernel(kind, m, v, outV, reps=25, flush_l2=False):
    """CUDA-event time (ms) of `reps` launches -- the timing loop of testSpMVImplCuda
    (test/SpMV_test.cu:103-145) with events instead of a host stopwatch."""
    t = np.zeros(reps, dtype=np.float32)
    check(lib().spmvb200_time_device(m.handle, kind, ptr(v), ptr(outV), reps, int(flush_l2), ptr(t)), "time_device")
    return t


def iterate(kind, m, a, b, iters, use_graph=True, stream=None):
    """x <- A x, `iters` times on one GPU, ping-pong between device vectors a (x on entry) and b; the result is in b if iters
    is odd, else in a.  Returns the CUDA-event time of all iterations in ms."""
    ms = C.c_float(0)
    check(lib().spmvb200_iterate_device(m.handle, kind, ptr(a), ptr(b), int(iters), int(bool(use_graph)), stream, C.byref(ms)), "iterate_device")
    return ms.value


def spmv_push(kind, m, v, outV, dst, lo, hi, row_offset, stream=None):
    """SpMV with fused output delivery: row r also goes to dst[p][r + row_offset] when lo[p] <= r + row_offset < hi[p]
    (dst: device pointers, e.g. the next x of the other GPUs mapped over CUDA IPC)."""
    p = capi.Push()
    p.n = len(dst)
    for i, (d, a_, b_) in enumerate(zip(dst, lo, hi)):
        p.dst[i], p.lo[i], p.hi[i] = ptr(d), int(a_), int(b_)
    p.row_offset = int(row_offset)
    check(lib().spmvb200_spmv_device_push(m.handle, kind, ptr(v), ptr(outV), C.byref(p), stream), "spmv_device_push")
    return EXIT_SUCCESS


def spmv_host(kind, m, x, y, timed=True):
    """x, y host buffers (numpy or pinned torch tensors): H2D x, kernel, D2H y.  Returns kernel ms; timed=False passes a NULL
    kernel_ms (no time-stamped events between the chunks of the pipelined path: ~10 % faster on long vectors) and returns None."""
    if not timed:
        check(lib().spmvb200_spmv_host(m.handle, kind, ptr(x), ptr(y), None), "spmv_host")
        return None
    ms = C.c_float(0)
    check(lib().spmvb200_spmv_host(m.handle, kind, ptr(x), ptr(y), C.byref(ms)), "spmv_host")
    return ms.value


# ---------------------------------------------------------------------------------------- SPMV_INTERF adapters
def _host_adapter(kind, is_ell):
    def f(mat, x, cfg, y):
        el = C.c_double(0)
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert isinstance(y, np.ndarray) and y.dtype == np.float64 and y.flags.c_contiguous and len(y) == mat.M
        rl = mat.RL
        check(lib().spmvb200_cached_spmv(id(mat), kind, int(is_ell), mat.M, mat.N, mat.MAX_ROW_NZ if is_ell else 0,
                                         ptr(mat.IRP), ptr(mat.JA), ptr(mat.AS), ptr(rl), ptr(x), ptr(y), C.byref(el)),
              "b200SpMV")
        f.ElapsedInternal = el.value
        return EXIT_SUCCESS
    f.ElapsedInternal = 0.0
    return f


b200SpMVRowsCSR = _host_adapter(CSR_ROWS, False)
b200SpMVWarpPerRowCSR = _host_adapter(CSR_ROWS_WARP, False)
b200SpMVAdaptiveCSR = _host_adapter(CSR_ADAPTIVE, False)
b200SpMVRowsSELL = _host_adapter(SELL_ROWS, False)
b200SpMVRowsXWIN = _host_adapter(XWIN_ROWS, False)
b200SpMVRowsELL = _host_adapter(ELL_ROWS, True)
b200SpMVRowsELLNNTransposed = _host_adapter(ELL_ROWS_NT, True)
b200SpMVWarpsPerRowELLNTrasposed = _host_adapter(ELL_ROWS_WARP_NT, True)
SpmvB200CSRFuncs = [b200SpMVRowsCSR, b200SpMVWarpPerRowCSR, b200SpMVAdaptiveCSR, b200SpMVRowsSELL, b200SpMVRowsXWIN]
SpmvB200ELLFuncs = [b200SpMVRowsELL, b200SpMVRowsELLNNTransposed, b200SpMVWarpsPerRowELLNTrasposed]


def cache_drop(mat=None):
    check(lib().spmvb200_cache_drop(None if mat is None else id(mat)), "cache_drop")


# ---------------------------------------------------------------------------------------- first-use picks, comparators
def tune(kind, m, v, outV, stream=None):
    """Make the first-use pick of a self-tuning kind now (blocking; outV is scratch).  Afterwards launches of that kind are pure
    asynchronous launches and can be captured into a CUDA graph with the tuned kernel inside (spmvb200_tune)."""
    check(lib().spmvb200_tune(m.handle, kind, ptr(v), ptr(outV), stream), "tune")
    return EXIT_SUCCESS


def set_tuning_mode(deterministic):
    """False (default): candidates are timed, the fastest wins (may differ from run to run).  True: the pick is a pure function of
    the matrix structure -- tolerance kinds then return the same bits in every process.  Process-wide."""
    check(lib().spmvb200_set_tuning_mode(1 if deterministic else 0), "set_tuning_mode")


def tuning_get(m):
    """the handle's picks as 8 integers (install them in another handle of the same matrix with tuning_set)"""
    p = (C.c_int32 * 8)()
    check(lib().spmvb200_tuning_get(m.handle, p), "tuning_get")
    return list(p)


def tuning_set(m, picks):
    p = (C.c_int32 * 8)(*[int(v) for v in picks])
    check(lib().spmvb200_tuning_set(m.handle, p), "tuning_set")


def compare_strict_csr(mat, x, y_ref, y, tau=1e-12):
    """|y_i - yref_i| <= tau * sum_j |a_ij x_j| for every row, NaN / Inf fails.  Returns (rows failing, worst ratio).
    Host arrays in, host loop inside the library: a comparator, not a compute path."""
    bad, worst = C.c_uint64(0), C.c_double(0)
    check(lib().spmvb200_compare_strict_csr(mat.M, ptr(mat.IRP), ptr(mat.JA), ptr(mat.AS), ptr(np.ascontiguousarray(x, dtype=np.float64)),
                                            ptr(np.ascontiguousarray(y_ref, dtype=np.float64)), ptr(np.ascontiguousarray(y, dtype=np.float64)),
                                            tau, C.byref(bad), C.byref(worst)), "compare_strict_csr")
    return bad.value, worst.value


def compare_abs(a, b, threshold=7e-4):
    """doubleVectorsDiff semantics (src/commons/utils.c:362-393; threshold DOUBLE_DIFF_THREASH = 7e-4), but NaN fails.
    Returns (failed, max |a-b|)."""
    a, b = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
    failed, dmax = C.c_int(0), C.c_double(0)
    check(lib().spmvb200_compare_abs(len(a), ptr(a), ptr(b), threshold, C.byref(failed), C.byref(dmax)), "compare_abs")
    return bool(failed.value), dmax.value
