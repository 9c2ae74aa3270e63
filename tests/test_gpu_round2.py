"""GPU tests of the round-2 surface of the C ABI: the row-block shard (one-call sharded step), capture-safe launches and the
explicit tune call, deterministic picks, the adapter cache's content fingerprint, structure validation at upload, the SELL
hybrid for skewed matrices, the chunked host path of the x-window kernel and in-place page-locking of pageable buffers."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TAU = 1e-12
STREAM_TILE = 2048


@pytest.fixture(scope="module")
def sp():
    import spmv_openmp_cuda_b200 as sp
    sp.capi.require_device()
    return sp


@pytest.fixture(scope="module")
def orc():
    import oracle
    return oracle


def _y(orc, mat, x):
    return orc.sgemv_serial(mat.IRP, mat.JA, mat.AS, x)


# ---------------------------------------------------------------------------------------------- shard (world = 1)
@pytest.mark.parametrize("kind_name", ["csr_rows", "xwin", "ell"])
def test_shard_single_rank_step_and_host(sp, orc, kind_name):
    """The sharded entry points with one rank: x <- A x iterates and the one-call host step equal the oracle bit for bit."""
    from spmv_openmp_cuda_b200.distributed import RowBlockShard
    s = sp.synth
    mat = s.host_csr(s.banded(150_000, 32, 3000))
    d_csr = sp.spMatCpyCSR(mat)
    if kind_name == "xwin":
        dm, kind = d_csr.to_xwin(512, 1024), sp.XWIN_ROWS
    elif kind_name == "ell":
        dm, kind = d_csr.to_ell(sp.FMT_ELL_COLMAJOR), sp.ELL_ROWS
    else:
        dm, kind = d_csr, sp.CSR_ROWS
    sh = RowBlockShard(dm, [0, mat.M], kind, nbuf=3, col_range=d_csr.col_range)
    x0 = s.host_vector(mat.N) * 1e3
    sh.set_x(0, x0)
    sh.step(0, 1)
    sh.step(1, 2)
    sh.step(2, 1)
    want = x0
    for _ in range(3):
        want = _y(orc, mat, want)
    np.testing.assert_array_equal(sh.rows_of(1), want)
    y_ref = _y(orc, mat, x0)
    y = np.full(mat.M, np.nan)
    for rep in range(3):  # pageable buffers: plain copies first, then page-locked in place by the explicit call
        y.fill(np.nan)
        if rep == 1:
            assert sp.capi.lib().spmvb200_host_register(sp.capi.ptr(x0), x0.nbytes) == 0
            assert sp.capi.lib().spmvb200_host_register(sp.capi.ptr(y), y.nbytes) == 0
        assert sh.spmv_host(x0, y) > 0
        np.testing.assert_array_equal(y, y_ref)
    y.fill(np.nan)
    assert sh.spmv_host(x0, y, timed=False) is None  # no time-stamped events between the chunks
    np.testing.assert_array_equal(y, y_ref)
    sh.close()
    assert sp.capi.lib().spmvb200_host_unregister(None) == 0


# ---------------------------------------------------------------------------------------------- capture / tune
def _graph_of(torch, fn):
    st = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        fn(torch.cuda.current_stream().cuda_stream)
    return g


@pytest.mark.parametrize("kind_name", ["CSR_ROWS", "CSR_ROWS_WARP", "CSR_ADAPTIVE", "ELL_ROWS"])
def test_untuned_handle_can_be_captured(sp, orc, kind_name):
    """First launch of a self-tuning kind INSIDE a stream capture: nothing is picked (no allocation, no synchronisation), the
    plain kernel is recorded, the graph replays correctly and the handle is still unpicked afterwards."""
    import torch
    s = sp.synth
    mat = s.host_csr(s.banded(300_000, 32, 2000))  # >= 2^20 nnz: tuning would build copies
    kind = getattr(sp, kind_name)
    d_csr = sp.spMatCpyCSR(mat)
    dm = d_csr.to_ell(sp.FMT_ELL_COLMAJOR) if kind == sp.ELL_ROWS else d_csr
    x = s.host_vector(mat.N)
    y_ref = _y(orc, mat, x)
    dx, dy = sp.DeviceVector.from_host(x), sp.DeviceVector(mat.M)
    lib = sp.capi.lib()
    picks = (C.c_int32 * 8)()
    g = _graph_of(torch, lambda st: sp.capi.check(lib.spmvb200_spmv_device(dm.handle, kind, dx.data_ptr(), dy.data_ptr(), st), "capture"))
    sp.capi.check(lib.spmvb200_tuning_get(dm.handle, picks), "tuning_get")
    if kind in (sp.CSR_ROWS, sp.ELL_ROWS):
        assert picks[1] == -1
    elif kind == sp.CSR_ADAPTIVE:
        assert picks[0] == -1
    else:
        assert picks[2] == 0
    for _ in range(2):
        dy.fill_bytes(0xFF)
        g.replay()
        torch.cuda.synchronize()
        y = dy.to_host()
        assert orc.strict_diff_csr(mat.IRP, mat.JA, mat.AS, x, y_ref, y, tau=TAU)[0] == 0
        if kind in (sp.CSR_ROWS, sp.ELL_ROWS):
            np.testing.assert_array_equal(y, y_ref)
    # explicit tune (blocking), then the TUNED kernel goes into a new graph
    sp.capi.check(lib.spmvb200_tune(dm.handle, kind, dx.data_ptr(), dy.data_ptr(), None), "tune")
    sp.capi.check(lib.spmvb200_tuning_get(dm.handle, picks), "tuning_get")
    assert (picks[1] >= 0) if kind in (sp.CSR_ROWS, sp.ELL_ROWS) else (picks[0] >= 0 if kind == sp.CSR_ADAPTIVE else picks[2] > 0)
    g2 = _graph_of(torch, lambda st: sp.capi.check(lib.spmvb200_spmv_device(dm.handle, kind, dx.data_ptr(), dy.data_ptr(), st), "capture"))
    dy.fill_bytes(0xFF)
    g2.replay()
    torch.cuda.synchronize()
    assert orc.strict_diff_csr(mat.IRP, mat.JA, mat.AS, x, y_ref, dy.to_host(), tau=TAU)[0] == 0


def test_deterministic_picks_and_pick_transfer(sp, orc):
    """Deterministic mode: two handles of one matrix pick the same kernels and the tolerance kinds return the same bits.
    Timed mode: picks read from one handle and installed in another give the same bits too."""
    s = sp.synth
    lib = sp.capi.lib()
    x = None
    try:
        for builder in (lambda: s.host_csr(s.banded(300_000, 32, 2000)), lambda: s.rmat_host_csr(15, 16), lambda: s.host_csr(s.mixed(300_000, 32, 0.1))):
            mat = builder()
            x = s.host_vector(mat.N)
            y_ref = _y(orc, mat, x)
            dx, dy = sp.DeviceVector.from_host(x), sp.DeviceVector(mat.M)
            sp.capi.check(lib.spmvb200_set_tuning_mode(1), "set_tuning_mode")
            assert lib.spmvb200_get_tuning_mode() == 1
            outs, picks = [], []
            for _ in range(2):
                dm = sp.spMatCpyCSR(mat)
                ys = []
                for f in (sp.cudaSpMVRowsCSR, sp.cudaSpMVWarpPerRowCSR, sp.cudaSpMVAdaptiveCSR):
                    dy.fill_bytes(0xFF)
                    f(dm, dx, sp.Config(), dy)
                    ys.append(dy.to_host())
                    assert orc.strict_diff_csr(mat.IRP, mat.JA, mat.AS, x, y_ref, ys[-1], tau=TAU)[0] == 0, f.__name__
                p = (C.c_int32 * 8)()
                sp.capi.check(lib.spmvb200_tuning_get(dm.handle, p), "tuning_get")
                outs.append(ys)
                picks.append(list(p))
                dm.free()
            assert picks[0] == picks[1]
            for a, b in zip(outs[0], outs[1]):
                np.testing.assert_array_equal(a, b)
            short = np.diff(mat.IRP) <= STREAM_TILE
            np.testing.assert_array_equal(outs[0][0][short], y_ref[short])
            # timed picks of handle A installed in handle B
            sp.capi.check(lib.spmvb200_set_tuning_mode(0), "set_tuning_mode")
            da, db = sp.spMatCpyCSR(mat), sp.spMatCpyCSR(mat)
            ya = []
            for f in (sp.cudaSpMVRowsCSR, sp.cudaSpMVWarpPerRowCSR, sp.cudaSpMVAdaptiveCSR):
                f(da, dx, sp.Config(), dy)
                ya.append(dy.to_host())
            p = (C.c_int32 * 8)()
            sp.capi.check(lib.spmvb200_tuning_get(da.handle, p), "tuning_get")
            sp.capi.check(lib.spmvb200_tuning_set(db.handle, p), "tuning_set")
            q = (C.c_int32 * 8)()
            sp.capi.check(lib.spmvb200_tuning_get(db.handle, q), "tuning_get")
            assert list(p) == list(q)
            for f, want in zip((sp.cudaSpMVRowsCSR, sp.cudaSpMVWarpPerRowCSR, sp.cudaSpMVAdaptiveCSR), ya):
                dy.fill_bytes(0xFF)
                f(db, dx, sp.Config(), dy)
                np.testing.assert_array_equal(dy.to_host(), want, err_msg=f.__name__)
    finally:
        lib.spmvb200_set_tuning_mode(0)


# ---------------------------------------------------------------------------------------------- adapter cache / validation
def test_adapter_cache_sees_a_changed_matrix(sp, orc):
    """Same key, same addresses, same shape -- different content: the cached device copy must not be reused."""
    s = sp.synth
    lib = sp.capi.lib()
    mat = s.host_csr(s.banded(20_000, 16, 500))
    x = s.host_vector(mat.N)
    y = np.empty(mat.M)
    key = 0x1234

    def call():
        sp.capi.check(lib.spmvb200_cached_spmv(key, sp.CSR_ROWS, 0, mat.M, mat.N, 0, sp.capi.ptr(mat.IRP), sp.capi.ptr(mat.JA), sp.capi.ptr(mat.AS),
                                               sp.capi.ptr(mat.RL), sp.capi.ptr(x), sp.capi.ptr(y), None), "cached_spmv")
        return y.copy()

    np.testing.assert_array_equal(call(), _y(orc, mat, x))
    mat.AS *= -2.0  # in place: a "new" matrix in the same buffers (what a freed-and-reallocated same-shape matrix looks like)
    np.testing.assert_array_equal(call(), _y(orc, mat, x))
    mat.JA[:] = (mat.JA + 1) % mat.N  # structure changes too (rows stay sorted except at the wrap, which CSR_ROWS does not need)
    np.testing.assert_array_equal(call(), _y(orc, mat, x))
    sp.capi.check(lib.spmvb200_cache_drop(key), "cache_drop")


def test_upload_rejects_malformed_structure(sp):
    s = sp.synth
    mat = s.host_csr(s.banded(5000, 8, 100))
    bad = sp.Spmat.csr(mat.N, mat.IRP, mat.JA.copy(), mat.AS)
    bad.JA[123] = mat.N + 7  # column id >= N
    with pytest.raises(sp.SpmvB200Error, match="column id"):
        sp.spMatCpyCSR(bad)
    irp = mat.IRP.copy()
    irp[100], irp[101] = irp[101], irp[100] - 1  # IRP[100] > IRP[101]
    bad2 = sp.Spmat(mat.M, mat.N, mat.NZ, mat.JA, mat.AS, IRP=irp, RL=mat.RL, MAX_ROW_NZ=mat.MAX_ROW_NZ)
    with pytest.raises(sp.SpmvB200Error, match="monotone"):
        sp.spMatCpyCSR(bad2)
    out = C.c_void_p()
    rc = sp.capi.lib().spmvb200_csr_upload(mat.M, mat.N, sp.capi.ptr(mat.IRP), sp.capi.ptr(mat.JA), None, 0, mat.M, C.byref(out))
    assert rc != 0 and b"null" in sp.capi.lib().spmvb200_last_error()
    ell = s.csr_to_ell_host(mat)
    ell.JA = ell.JA.copy()  # (rows of equal length: the ELL arrays alias the CSR arrays)
    ell.JA[5] = mat.N  # first row, valid slot
    with pytest.raises(sp.SpmvB200Error, match="column id"):
        sp.spMatCpyELL(ell)
    sp.spMatCpyCSR(mat).free()  # and the engine is still usable after the rejections


# ---------------------------------------------------------------------------------------------- SELL on skewed rows
def test_sell_rows_on_skewed_matrix_hands_long_rows_over(sp, orc):
    """A stand-alone SELL handle of a power-law matrix: rows longer than 256 do not go through one thread each (round 1: 37 ms on
    R-MAT scale 22) but through the per-row kernels; results bit-identical up to 2048-entry rows, within tolerance beyond."""
    s = sp.synth
    mat = s.rmat_host_csr(17, 16)
    lens = np.diff(mat.IRP)
    assert lens.max() > 2048
    x = s.host_vector(mat.N)
    y_ref = _y(orc, mat, x)
    d_csr = sp.spMatCpyCSR(mat)
    d_sell = d_csr.to_sell()
    d_csr.free()  # the SELL handle keeps its own copy of the long rows
    dx, dy = sp.DeviceVector.from_host(x), sp.DeviceVector(mat.M)
    for _ in range(2):
        dy.fill_bytes(0xFF)
        sp.cudaSpMVRowsSELL(d_sell, dx, sp.Config(), dy)
        y = dy.to_host()
        assert orc.strict_diff_csr(mat.IRP, mat.JA, mat.AS, x, y_ref, y, tau=TAU)[0] == 0
        np.testing.assert_array_equal(y[lens <= STREAM_TILE], y_ref[lens <= STREAM_TILE])
    t = sp.time_kernel(sp.SELL_ROWS, d_sell, dx, dy, reps=5)
    t_csr = sp.time_kernel(sp.CSR_ROWS_WARP, sp.spMatCpyCSR(mat), dx, dy, reps=5)
    assert np.min(t) < 8 * np.min(t_csr), (t, t_csr)  # same league as the sub-warp kernel, not 100x off
    # the host adapter (CSR in, SELL built on the device) takes the same route
    yh = np.empty(mat.M)
    sp.b200SpMVRowsSELL(mat, x, sp.Config(), yh)
    np.testing.assert_array_equal(yh, y)
    sp.cache_drop()


# ---------------------------------------------------------------------------------------------- host path
@pytest.mark.parametrize("pinned", [False, True])
def test_host_path_xwindow_runs_in_row_block_chunks(sp, orc, pinned, monkeypatch):
    """spmvb200_spmv_host on an x-window handle / an exact kind that picked the x-window copy: row-block chunks, x pieces bounded
    by the windows the chunks read.  Pageable and page-locked buffers, several chunk counts."""
    import torch
    s = sp.synth
    mat = s.host_csr(s.banded(400_000, 32, 6000))
    x = s.host_vector(mat.N)
    y_ref = _y(orc, mat, x)
    d_csr = sp.spMatCpyCSR(mat)
    d_xw = d_csr.to_xwin()

    def buf(n):
        return torch.empty(n, dtype=torch.float64).pin_memory().numpy() if pinned else np.empty(n)
    xb = buf(mat.N)
    xb[:] = x
    for chunks in ("1", "3", "7"):
        monkeypatch.setenv("SPMVB200_HOST_CHUNKS", chunks)
        for dm, kind in ((d_xw, sp.XWIN_ROWS), (d_csr, sp.CSR_ROWS)):
            y = buf(mat.M)
            for rep in range(4):
                y.fill(np.nan)
                if rep == 1 and not pinned:  # pageable -> page-locked in place (explicit: the library never guesses a buffer's lifetime)
                    assert sp.capi.lib().spmvb200_host_register(sp.capi.ptr(xb), xb.nbytes) == 0
                    assert sp.capi.lib().spmvb200_host_register(sp.capi.ptr(y), y.nbytes) == 0
                if rep == 3:
                    assert sp.spmv_host(kind, dm, xb, y, timed=False) is None
                else:
                    assert sp.spmv_host(kind, dm, xb, y) > 0
                np.testing.assert_array_equal(y, y_ref, err_msg="%s chunks=%s rep=%d" % (kind, chunks, rep))
            assert sp.capi.lib().spmvb200_host_unregister(sp.capi.ptr(y)) == 0  # before y is freed
    assert sp.capi.lib().spmvb200_host_unregister(None) == 0


def test_host_register_auto_policy_is_opt_in(sp):
    """Without SPMVB200_HOST_REGISTER=auto the library leaves caller buffers alone: a pageable buffer stays pageable however often
    it comes back (a buffer freed while registered would poison its address range for every later CUDA call)."""
    import os
    if os.environ.get("SPMVB200_HOST_REGISTER") == "auto":
        pytest.skip("this process opted in to the automatic policy")
    s = sp.synth
    mat = s.host_csr(s.banded(300_000, 8, 500))
    dm = sp.spMatCpyCSR(mat)
    x, y = s.host_vector(mat.N), np.empty(mat.M)
    L = sp.capi.lib()
    for _ in range(3):
        sp.spmv_host(sp.CSR_ROWS, dm, x, y)
    assert L.spmvb200_host_registered(sp.capi.ptr(x)) == 0 and L.spmvb200_host_registered(sp.capi.ptr(y)) == 0
    assert L.spmvb200_host_register(sp.capi.ptr(x), x.nbytes) == 0
    assert L.spmvb200_host_register(sp.capi.ptr(x), x.nbytes) == 0  # idempotent
    assert L.spmvb200_host_registered(sp.capi.ptr(x)) == 1
    sp.spmv_host(sp.CSR_ROWS, dm, x, y)
    assert L.spmvb200_host_unregister(sp.capi.ptr(x)) == 0
    assert L.spmvb200_host_registered(sp.capi.ptr(x)) == 0
    assert sp.capi.lib().spmvb200_host_unregister(sp.capi.ptr(x)) == 0  # unknown pointer: no-op


# ---------------------------------------------------------------------------------------------- x-window launch shapes
@pytest.mark.parametrize("R,W,w", [(512, 1024, 3000), (1024, 2048, 9000), (512, 256, 300)])
def test_xwin_launch_shapes_agree_bit_for_bit(sp, orc, R, W, w):
    """One CTA per row block (0) and persistent CTAs over contiguous row blocks (1: window ring and prefetch carried across row
    blocks) return the oracle's bits; so do the row-block chunks of the host path and the delivery epilogue of the shard step."""
    s = sp.synth
    mat = s.host_csr(s.banded(400_003, 32, w))  # > 148 row blocks for every R here, last row block ragged
    x = s.host_vector(mat.N) * 1e3
    y_ref = _y(orc, mat, x)
    d_csr = sp.spMatCpyCSR(mat)
    dxw = d_csr.to_xwin(R, W)
    dx, dy = sp.DeviceVector.from_host(x), sp.DeviceVector(mat.M)
    for shape in (0, 1):
        sp.tuning_set(dxw, [-1, -1, 0, shape, -1, 0, 0, 0])
        assert sp.tuning_get(dxw)[3] == shape
        dy.fill_bytes(0xFF)
        sp.cudaSpMVRowsXWIN(dxw, dx, sp.Config(), dy)
        np.testing.assert_array_equal(dy.to_host(), y_ref)
        yh = np.full(mat.M, np.nan)
        sp.spmv_host(sp.XWIN_ROWS, dxw, x, yh, timed=False)
        np.testing.assert_array_equal(yh, y_ref)
        from spmv_openmp_cuda_b200.distributed import RowBlockShard
        sh = RowBlockShard(dxw, [0, mat.M], sp.XWIN_ROWS, nbuf=2, col_range=d_csr.col_range)
        sh.set_x(0, x)
        sh.step(0, 1)
        np.testing.assert_array_equal(sh.rows_of(1), y_ref)
        sh.close()
