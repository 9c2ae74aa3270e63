"""ctypes binding of include/spmv_b200.h (libspmv_b200.so).  No CPU fallback: if the CUDA library is
missing or a call fails, an exception is raised -- nothing here computes on the host."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libspmv_b200.so")

# kernel selectors / formats (include/spmv_b200.h)
CSR_ROWS, CSR_ROWS_WARP, ELL_ROWS, ELL_ROWS_NT, ELL_ROWS_WARP_NT, CSR_ADAPTIVE, SELL_ROWS, XWIN_ROWS = range(8)
FMT_CSR, FMT_ELL_COLMAJOR, FMT_ELL_ROWMAJOR, FMT_SELL, FMT_XWIN = range(5)


class SpmvB200Error(RuntimeError):
    pass


class Push(C.Structure):
    """spmvb200_push: extra destinations of an SpMV's output rows (include/spmv_b200.h)."""
    _fields_ = [("n", C.c_int), ("dst", C.c_void_p * 8), ("lo", C.c_uint64 * 8), ("hi", C.c_uint64 * 8), ("row_offset", C.c_uint64)]


class Synth(C.Structure):
    _fields_ = [("kind", C.c_int), ("seed", C.c_uint64), ("p0", C.c_uint64), ("p1", C.c_uint64), ("p2", C.c_uint64),
                ("p3", C.c_uint64)]


_u64 = C.c_uint64
_vp = C.c_void_p
_SIGS = {
    "spmvb200_last_error": (C.c_char_p, []),
    "spmvb200_version": (C.c_int, []),
    "spmvb200_launch_count": (C.c_ulonglong, []),
    "spmvb200_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "spmvb200_set_device": (C.c_int, [C.c_int]),
    "spmvb200_device_info": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "spmvb200_csr_upload": (C.c_int, [_u64, _u64, _vp, _vp, _vp, _u64, _u64, C.POINTER(_vp)]),
    "spmvb200_ell_upload": (C.c_int, [_u64, _u64, _u64, _vp, _vp, _vp, _u64, _u64, C.c_int, C.POINTER(_vp)]),
    "spmvb200_csr_adopt_device": (C.c_int, [_u64, _u64, _u64, _vp, _vp, _vp, C.c_int, C.POINTER(_vp)]),
    "spmvb200_ell_from_csr": (C.c_int, [_vp, C.c_int, C.POINTER(_vp)]),
    "spmvb200_sell_from_csr": (C.c_int, [_vp, C.c_uint32, C.POINTER(_vp)]),
    "spmvb200_xwin_from_csr": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.POINTER(_vp)]),
    "spmvb200_xwin_info": (C.c_int, [_vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(_u64)]),
    "spmvb200_free": (C.c_int, [_vp]),
    "spmvb200_dims": (C.c_int, [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64), C.POINTER(C.c_int)]),
    "spmvb200_algorithmic_bytes": (_u64, [_vp]),
    "spmvb200_device_bytes": (_u64, [_vp]),
    "spmvb200_index_bits": (C.c_int, [_vp]),
    "spmvb200_kind_supported": (C.c_int, [_vp, C.c_int]),
    "spmvb200_kind_name": (C.c_char_p, [C.c_int]),
    "spmvb200_adaptive_choice": (C.c_int, [_vp, C.c_char_p, C.c_size_t]),
    "spmvb200_exact_choice": (C.c_int, [_vp, C.c_char_p, C.c_size_t]),
    "spmvb200_spmv_device": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp]),
    "spmvb200_spmv_device_push": (C.c_int, [_vp, C.c_int, _vp, _vp, C.POINTER(Push), _vp]),
    "spmvb200_tune": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp]),
    "spmvb200_set_tuning_mode": (C.c_int, [C.c_int]),
    "spmvb200_get_tuning_mode": (C.c_int, []),
    "spmvb200_tuning_get": (C.c_int, [_vp, C.POINTER(C.c_int32)]),
    "spmvb200_tuning_set": (C.c_int, [_vp, C.POINTER(C.c_int32)]),
    "spmvb200_shard_create": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(_u64), C.c_int, C.POINTER(_u64), C.POINTER(_vp)]),
    "spmvb200_shard_blob_bytes": (C.c_size_t, [_vp]),
    "spmvb200_shard_export": (C.c_int, [_vp, _vp]),
    "spmvb200_shard_connect": (C.c_int, [_vp, _vp]),
    "spmvb200_shard_x": (_vp, [_vp, C.c_int]),
    "spmvb200_shard_halo_rows": (C.c_int, [_vp, C.POINTER(_u64)]),
    "spmvb200_shard_step": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "spmvb200_shard_barrier": (C.c_int, [_vp, _vp]),
    "spmvb200_shard_spmv_host": (C.c_int, [_vp, _vp, _vp, C.POINTER(C.c_float)]),
    "spmvb200_shard_free": (C.c_int, [_vp]),
    "spmvb200_host_unregister": (C.c_int, [_vp]),
    "spmvb200_host_register": (C.c_int, [_vp, C.c_size_t]),
    "spmvb200_host_registered": (C.c_int, [_vp]),
    "spmvb200_host_alloc": (_vp, [C.c_size_t]),
    "spmvb200_host_free": (C.c_int, [_vp]),
    "spmvb200_compare_strict_csr": (C.c_int, [_u64, _vp, _vp, _vp, _vp, _vp, _vp, C.c_double, C.POINTER(_u64), C.POINTER(C.c_double)]),
    "spmvb200_compare_abs": (C.c_int, [_u64, _vp, _vp, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "spmvb200_ipc_export": (C.c_int, [_vp, _vp]),
    "spmvb200_ipc_open": (C.c_int, [_vp, C.POINTER(_vp)]),
    "spmvb200_ipc_close": (C.c_int, [_vp]),
    "spmvb200_peer_barrier": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_uint32, _vp]),
    "spmvb200_iterate_device": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp, C.POINTER(C.c_float)]),
    "spmvb200_spmv_host": (C.c_int, [_vp, C.c_int, _vp, _vp, C.POINTER(C.c_float)]),
    "spmvb200_time_device": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp]),
    "spmvb200_cached_spmv": (C.c_int, [_vp, C.c_int, C.c_int, _u64, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(C.c_double)]),
    "spmvb200_cache_drop": (C.c_int, [_vp]),
    "spmvb200_dmalloc": (C.c_int, [C.POINTER(_vp), C.c_size_t]),
    "spmvb200_dfree": (C.c_int, [_vp]),
    "spmvb200_h2d": (C.c_int, [_vp, _vp, C.c_size_t]),
    "spmvb200_d2h": (C.c_int, [_vp, _vp, C.c_size_t]),
    "spmvb200_sync": (C.c_int, []),
    "spmvb200_h2d_async": (C.c_int, [_vp, _vp, C.c_size_t, _vp]),
    "spmvb200_d2h_async": (C.c_int, [_vp, _vp, C.c_size_t, _vp]),
    "spmvb200_stream_sync": (C.c_int, [_vp]),
    "spmvb200_push_rows": (C.c_int, [_vp, _u64, C.POINTER(Push), _vp]),
    "spmvb200_synth_dims": (C.c_int, [C.POINTER(Synth), C.POINTER(_u64), C.POINTER(_u64)]),
    "spmvb200_synth_rowlen_host": (C.c_int, [C.POINTER(Synth), _u64, _u64, _vp]),
    "spmvb200_synth_fill_host": (C.c_int, [C.POINTER(Synth), _u64, _u64, _vp, _vp, _vp]),
    "spmvb200_synth_csr_device": (C.c_int, [C.POINTER(Synth), _u64, _u64, C.POINTER(_vp)]),
    "spmvb200_synth_rmat_keys_host": (C.c_int, [C.c_int, _u64, _u64, _u64, _vp]),
    "spmvb200_synth_rmat_values_host": (C.c_int, [_u64, _u64, _vp, _vp]),
    "spmvb200_synth_rmat_csr_device": (C.c_int, [C.c_int, _u64, _u64, C.POINTER(_vp)]),
    "spmvb200_synth_vector_host": (C.c_int, [_u64, _u64, _u64, C.c_double, _vp]),
    "spmvb200_synth_vector_device": (C.c_int, [_u64, _u64, _u64, C.c_double, _vp]),
    "spmvb200_col_range": (C.c_int, [_vp, C.POINTER(_u64), C.POINTER(_u64)]),
    "spmvb200_csr_download": (C.c_int, [_vp, _vp, _vp, _vp]),
}
EXPORTS = tuple(sorted(_SIGS))

_lib = None


def lib():
    """Load libspmv_b200.so (building it in-tree if nvcc is available and it is missing/stale)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            try:
                from . import build as _b
                _b.build()
            except Exception as e:  # noqa: BLE001
                raise SpmvB200Error("CUDA extension %s is missing and could not be built (%s); "
                                    "there is no CPU fallback" % (LIB_PATH, e)) from e
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().spmvb200_last_error().decode(errors="replace")
        raise SpmvB200Error("%s failed: %s" % (what or "spmv_b200 call", msg))


def ptr(a):
    """Raw address of a numpy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError("cannot take the address of %r" % type(a))


def device_count():
    n = C.c_int(0)
    rc = lib().spmvb200_device_count(C.byref(n))
    return 0 if rc else n.value


def require_device():
    if device_count() < 1:
        raise SpmvB200Error("no CUDA device visible: the B200 SpMV engine has no CPU fallback")


def device_info():
    name = C.create_string_buffer(128)
    sm, l2, mem = C.c_int(0), C.c_size_t(0), C.c_size_t(0)
    check(lib().spmvb200_device_info(name, 128, C.byref(sm), C.byref(l2), C.byref(mem)), "device_info")
    return dict(name=name.value.decode(), sm_count=sm.value, l2_bytes=l2.value, mem_bytes=mem.value)
