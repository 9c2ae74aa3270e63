"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle / golden fixtures.

Tolerances (stated, as BASELINE.json's north_star asks):
  * cudaSpMVRowsCSR and cudaSpMVRowsELL replacements sum each row left to right with separate
    mul/add roundings => BIT-IDENTICAL to sgemvSerial (src/SpMV_CSR_OMP.c:229-250);
  * tree-reduced kinds: |y_i - yref_i| <= 1e-12 * sum_j |a_ij x_j|   (TAU), NaN => fail;
  * every kind also passes the reference's own check |y - yref| <= 7e-4 (src/commons/utils.c:362-393).
Output vectors are pre-filled with a NaN pattern so stale results can never pass.
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu
TAU = 1e-12


@pytest.fixture(scope="module")
def sp():
    import spmv_openmp_cuda_b200 as sp
    sp.capi.require_device()
    return sp


@pytest.fixture(scope="module")
def orc():
    import oracle
    return oracle


EXACT = {"cudaSpMVRowsCSR", "cudaSpMVRowsELL", "cudaSpMVRowsSELL", "cudaSpMVRowsXWIN"}
STREAM_TILE = 2048  # csrc/kernels.cuh


def run_all_kinds(sp, orc, mat, x, y_ref, ell=None, kinds="all"):
    """The reference harness flow (test/SpMV_test.cu:273-343): upload, run every table entry, compare."""
    cfg = sp.Config()
    dx = sp.DeviceVector.from_host(x)
    dy = sp.DeviceVector(mat.M)
    ell = ell if ell is not None else sp.synth.csr_to_ell_host(mat)
    d_csr = sp.spMatCpyCSR(mat)
    d_ell = sp.spMatCpyELL(ell)
    d_rm = sp.spMatCpyELLNNPitched(ell)
    d_sell = d_csr.to_sell(64)       # tiny sorting window: slices straddle every length class
    d_sell2 = d_csr.to_sell()        # default window
    # x-window CSR: 64-column windows never overflow the 255-per-(row, window) limit and give many tiles per row block;
    # the default geometry (4096 x 8192) applies when no row is longer than 255
    d_xw = [d_csr.to_xwin(512, 64), d_csr.to_xwin(1024, 200)] if mat.M else []
    if mat.M and mat.MAX_ROW_NZ <= 255:
        d_xw.append(d_csr.to_xwin())
    runs = [(f, d_csr) for f in sp.SpmvCUDA_CSRFuncs] + \
           [(sp.cudaSpMVRowsELL, d_ell), (sp.cudaSpMVRowsELLNNTransposed, d_rm), (sp.cudaSpMVWarpsPerRowELLNTrasposed, d_rm),
            (sp.cudaSpMVRowsSELL, d_sell), (sp.cudaSpMVRowsSELL, d_sell2)] + [(sp.cudaSpMVRowsXWIN, d) for d in d_xw]
    sorted_rows = all(np.all(np.diff(mat.JA[int(a):int(b)].astype(np.int64)) >= 0) for a, b in zip(mat.IRP[:-1], mat.IRP[1:])) \
        if mat.M <= 5000 else True  # generators emit sorted rows; fixtures are small enough to check
    for f, dm in runs:
        dy.fill_bytes(0xFF)
        assert f(dm, dx, cfg, dy) == 0
        y = dy.to_host()
        assert np.all(np.isfinite(y)), f.__name__
        failed, dmax = orc.double_vectors_diff(y_ref, y)
        assert not failed, (f.__name__, dmax)
        bad, worst = orc.strict_diff_csr(mat.IRP, mat.JA, mat.AS, x, y_ref, y, tau=TAU)
        assert bad == 0, (f.__name__, bad, worst)
        if f.__name__ in EXACT:
            # rows longer than one tile (2048 nnz) are split across CTAs by the CSR kernel and summed
            # in segment order: deterministic, within TAU, but not the serial order
            # (a stand-alone SELL handle hands its rows longer than 256 to the same per-row / per-segment kernels)
            exact = np.ones(mat.M, dtype=bool) if f.__name__ not in ("cudaSpMVRowsCSR", "cudaSpMVRowsSELL") else (np.diff(mat.IRP) <= STREAM_TILE)
            if f.__name__ == "cudaSpMVRowsXWIN" and not sorted_rows:
                continue  # windows are visited in column order: bit-identical only for column-sorted rows
            np.testing.assert_array_equal(y[exact], y_ref[exact], err_msg=f.__name__)
    for dm in [d_csr, d_ell, d_rm, d_sell, d_sell2] + d_xw:
        sp.cudaFreeSpmat(dm)


@pytest.mark.parametrize("name", golden_names())
def test_golden_fixtures_all_kinds(sp, orc, name):
    g = load_golden(name)
    mat = sp.Spmat.csr(int(g["N"]), g["irp"], g["ja"], g["as_"], RL=g["rl"])
    ell = sp.Spmat.ell(int(g["M"]), int(g["N"]), int(g["K"]), g["ell_ja"], g["ell_as"], RL=g["rl"], NZ=int(g["NZ"]))
    run_all_kinds(sp, orc, mat, g["x"], g["y_sgemvSerial"], ell=ell)


@pytest.mark.parametrize("name", ["lap2d_33", "rect_empty_rows", "skew_long_row"])
def test_ell_without_rowlens(sp, orc, name):
    """A driver built without -DROWLENS has no RL vector: lengths are derived on the device."""
    g = load_golden(name)
    ell = sp.Spmat.ell(int(g["M"]), int(g["N"]), int(g["K"]), g["ell_ja"], g["ell_as"], RL=None, NZ=int(g["NZ"]))
    dx, dy = sp.DeviceVector.from_host(g["x"]), sp.DeviceVector(int(g["M"]))
    for up, f in ((sp.spMatCpyELL, sp.cudaSpMVRowsELL), (sp.spMatCpyELLNNPitched, sp.cudaSpMVWarpsPerRowELLNTrasposed)):
        dm = up(ell, use_rowlens=False)
        dy.fill_bytes(0xFF)
        f(dm, dx, sp.Config(), dy)
        y = dy.to_host()
        bad, worst = orc.strict_diff_csr(g["irp"], g["ja"], g["as_"], g["x"], g["y_sgemvSerial"], y, tau=TAU)
        assert bad == 0, worst


def _oracle_y(orc, mat, x):
    return orc.sgemv_serial(mat.IRP, mat.JA, mat.AS, x)


@pytest.mark.parametrize("builder", [
    lambda s: s.host_csr(s.lap2d(64)),
    lambda s: s.host_csr(s.stencil27(20, 18, 16)),
    lambda s: s.host_csr(s.banded(30000, 32, 2000)),
    lambda s: s.host_csr(s.mixed(40000, 48, 0.03)),
    lambda s: s.rmat_host_csr(13, 16),
])
def test_scaled_down_baseline_configs(sp, orc, builder):
    mat = builder(sp.synth)
    x = sp.synth.host_vector(mat.N)
    run_all_kinds(sp, orc, mat, x, _oracle_y(orc, mat, x))


def test_edge_cases(sp, orc):
    rng = np.random.default_rng(7)

    def rand_csr(M, N, lens):
        irp = np.zeros(M + 1, dtype=np.uint64)
        irp[1:] = np.cumsum(lens)
        ja = np.concatenate([np.sort(rng.choice(N, k, replace=False)) for k in lens] + [np.zeros(0, int)]).astype(np.uint64)
        return sp.Spmat.csr(N, irp, ja, rng.uniform(-1, 1, int(irp[-1])))

    cases = {
        "all_rows_empty": rand_csr(100, 50, np.zeros(100, int)),
        "one_by_one": rand_csr(1, 1, np.array([1])),
        "rows_longer_than_a_tile": rand_csr(6, 20000, np.array([3, 9000, 0, 2048, 2049, 17000])),
        "many_short_then_long": rand_csr(3000, 6000, np.r_[np.ones(2999, int), 5000]),
        "tile_boundary_rows": rand_csr(1200, 4096, np.r_[np.full(400, 5), np.full(400, 0), np.full(400, 31)]),
        "single_column": rand_csr(777, 1, rng.integers(0, 2, 777)),
    }
    for name, mat in cases.items():
        x = rng.uniform(-1, 1, mat.N)
        y_ref = _oracle_y(orc, mat, x)
        if mat.MAX_ROW_NZ * mat.M < 3_000_000:
            run_all_kinds(sp, orc, mat, x, y_ref)
        else:  # ELL would be huge (the reference caps it too, config.h:69): CSR kinds only
            dx, dy, dm = sp.DeviceVector.from_host(x), sp.DeviceVector(mat.M), sp.spMatCpyCSR(mat)
            dsell = dm.to_sell(32)
            dxw = dm.to_xwin(512, 128)
            for f in sp.SpmvCUDA_CSRFuncs + [sp.cudaSpMVRowsSELL, sp.cudaSpMVRowsXWIN]:
                dy.fill_bytes(0xFF)
                f(dsell if f is sp.cudaSpMVRowsSELL else dxw if f is sp.cudaSpMVRowsXWIN else dm, dx, sp.Config(), dy)
                bad, worst = orc.strict_diff_csr(mat.IRP, mat.JA, mat.AS, x, y_ref, dy.to_host(), tau=TAU)
                assert bad == 0, (name, f.__name__, worst)


def test_unsorted_columns_all_kinds(sp, orc):
    """The reference's parser keeps the file order inside a row (src/lib/parser.c:157-215): rows need not be column-sorted.
    Every kind must still be right (the x-window build walks such rows instead of bisecting them; its windows are visited in
    column order, so it meets the tolerance rather than bit-exactness there)."""
    rng = np.random.default_rng(11)
    M, N = 5000, 40000
    lens = rng.integers(0, 40, M)
    irp = np.zeros(M + 1, dtype=np.uint64)
    irp[1:] = np.cumsum(lens)
    ja = np.concatenate([rng.permutation(rng.choice(N, k, replace=False)) for k in lens]).astype(np.uint64)
    mat = sp.Spmat.csr(N, irp, ja, rng.uniform(-1, 1, int(irp[-1])))
    x = rng.uniform(-1, 1, N)
    run_all_kinds(sp, orc, mat, x, _oracle_y(orc, mat, x))


def test_xwin_limits_fail_loudly(sp):
    """More than 255 non-zeros of one row inside one window, or a bad geometry: an error, never a wrong answer."""
    irp = np.array([0, 300, 301], dtype=np.uint64)
    ja = np.r_[np.arange(300), 5].astype(np.uint64)
    dm = sp.spMatCpyCSR(sp.Spmat.csr(1000, irp, ja, np.ones(301)))
    with pytest.raises(sp.SpmvB200Error, match="255"):
        dm.to_xwin(512, 1000)
    for R, W in ((500, 64), (512, 63), (16384, 64), (512, 70000)):
        with pytest.raises(sp.SpmvB200Error):
            dm.to_xwin(R, W)
    ok = dm.to_xwin(512, 128)  # 128-column windows: at most 128 per (row, window)
    dx, dy = sp.DeviceVector.from_host(np.ones(1000)), sp.DeviceVector(2)
    sp.cudaSpMVRowsXWIN(ok, dx, sp.Config(), dy)
    np.testing.assert_array_equal(dy.to_host(), [300.0, 1.0])
    with pytest.raises(sp.SpmvB200Error):
        sp.cudaSpMVRowsXWIN(dm, dx, sp.Config(), dy)  # CSR handle, x-window kind


def test_xwin_unaligned_x_and_odd_windows(sp, orc):
    """x that is not 16-byte aligned (the producer warp copies the windows itself) and N that leaves an odd last window."""
    mat = sp.synth.host_csr(sp.synth.banded(20001, 32, 700))
    x = sp.synth.host_vector(mat.N)
    y_ref = _oracle_y(orc, mat, x)
    dm = sp.spMatCpyCSR(mat)
    big = sp.DeviceVector(mat.N + 1)
    sp.capi.check(sp.capi.lib().spmvb200_h2d(big.data_ptr() + 8, sp.capi.ptr(x), x.nbytes), "h2d")
    dy = sp.DeviceVector(mat.M)
    for R, W in ((512, 1000), (1024, 4096), (4096, 8192), (2048, 350)):
        dxw = dm.to_xwin(R, W)
        for xp in (big.data_ptr() + 8, sp.DeviceVector.from_host(x)):
            dy.fill_bytes(0xFF)
            sp.cudaSpMVRowsXWIN(dxw, xp, sp.Config(), dy)
            np.testing.assert_array_equal(dy.to_host(), y_ref)
        dxw.free()


@pytest.mark.parametrize("cand,name", [(12, "xwindow"), (13, "sell")])
def test_adaptive_child_formats(sp, orc, cand, name, monkeypatch):
    """The adaptive mode's tuning run may build an x-window or SELL copy of the matrix; force each and check the result."""
    monkeypatch.setenv("SPMVB200_FORCE_CAND", str(cand))
    if sp.capi.lib().spmvb200_get_tuning_mode() != 0:
        pytest.skip("candidates are forced through the TIMED first-use pick; this process runs with SPMVB200_TUNE=deterministic")
    mat = sp.synth.host_csr(sp.synth.banded(70000, 32, 20000))  # 2.2 M nnz: above the size threshold for child formats
    x = sp.synth.host_vector(mat.N)
    y_ref = _oracle_y(orc, mat, x)
    dm, dx, dy = sp.spMatCpyCSR(mat), sp.DeviceVector.from_host(x), sp.DeviceVector(mat.M)
    for _ in range(3):
        dy.fill_bytes(0xFF)
        sp.cudaSpMVAdaptiveCSR(dm, dx, sp.Config(), dy)
        np.testing.assert_array_equal(dy.to_host(), y_ref)  # both child kernels sum in the serial order
    assert dm.adaptive_choice == name
    dy.fill_bytes(0xFF)
    sp.cudaSpMVRowsCSR(dm, dx, sp.Config(), dy)  # the CSR kinds still work on the same handle
    np.testing.assert_array_equal(dy.to_host(), y_ref)


def test_iterated_spmv_graph_and_plain(sp, orc):
    """x <- A x repeated: CUDA-graph replay and plain launches give the oracle's iterates bit for bit."""
    mat = sp.synth.host_csr(sp.synth.lap2d(96))
    x0 = sp.synth.host_vector(mat.N) * 1e4
    refs = [x0]
    for _ in range(7):
        refs.append(orc.sgemv_serial(mat.IRP, mat.JA, mat.AS, refs[-1]))
    d_csr = sp.spMatCpyCSR(mat)
    handles = [(d_csr, sp.CSR_ROWS), (d_csr.to_ell(sp.FMT_ELL_COLMAJOR), sp.ELL_ROWS), (d_csr.to_xwin(512, 512), sp.XWIN_ROWS),
               (d_csr.to_sell(64), sp.SELL_ROWS)]
    for dm, kind in handles:
        for iters in (0, 1, 6, 7):
            for graph in (True, False):
                a, b = sp.DeviceVector.from_host(x0), sp.DeviceVector(mat.N)
                b.fill_bytes(0xFF)
                ms = sp.iterate(kind, dm, a, b, iters, use_graph=graph)
                got = (b if iters % 2 else a).to_host()
                np.testing.assert_array_equal(got, refs[iters], err_msg="kind %d iters %d graph %s" % (kind, iters, graph))
                assert ms >= 0
    rect = sp.spMatCpyCSR(sp.Spmat.csr(7, np.array([0, 1, 2], dtype=np.uint64), np.array([0, 6], dtype=np.uint64), np.ones(2)))
    with pytest.raises(sp.SpmvB200Error, match="square"):
        sp.iterate(sp.CSR_ROWS, rect, sp.DeviceVector(7), sp.DeviceVector(7), 2)


def test_fused_output_delivery(sp, orc):
    """spmvb200_spmv_device_push: rows inside a destination's range land there (at their global index), nothing else is
    touched; fused epilogue (ELL, x-window) and the extra pass (other kinds) agree."""
    mat = sp.synth.host_csr(sp.synth.banded(9000, 32, 700))
    x = sp.synth.host_vector(mat.N)
    y_ref = _oracle_y(orc, mat, x)
    a, b = 2000, 7000  # this "rank" owns rows [a, b) of a larger job
    d_csr = sp.spMatCpyCSR(mat, a, b)
    handles = [(d_csr, sp.CSR_ROWS), (d_csr, sp.CSR_ROWS_WARP), (d_csr, sp.CSR_ADAPTIVE), (d_csr.to_ell(sp.FMT_ELL_COLMAJOR), sp.ELL_ROWS),
               (d_csr.to_xwin(512, 1024), sp.XWIN_ROWS), (d_csr.to_xwin(1024, 8192), sp.XWIN_ROWS)]
    dx = sp.DeviceVector.from_host(x)
    ranges = [(1500, 2600), (6990, 9000), (3000, 3001)]
    for dm, kind in handles:
        dy = sp.DeviceVector(b - a)
        dsts = [sp.DeviceVector(mat.M) for _ in ranges]
        for d in dsts + [dy]:
            d.fill_bytes(0xFF)
        sp.spmv_push(kind, dm, dx, dy, dsts, [r[0] for r in ranges], [r[1] for r in ranges], a)
        sp.capi.check(sp.capi.lib().spmvb200_sync(), "sync")
        exact = kind in (sp.CSR_ROWS, sp.ELL_ROWS, sp.XWIN_ROWS)
        y = dy.to_host()
        (np.testing.assert_array_equal if exact else np.testing.assert_allclose)(y, y_ref[a:b])
        for d, (lo, hi) in zip(dsts, ranges):
            got = d.to_host()
            lo2, hi2 = max(lo, a), min(hi, b)
            np.testing.assert_array_equal(got[lo2:hi2], y[lo2 - a:hi2 - a])
            assert np.all(np.isnan(got[:lo2])) and np.all(np.isnan(got[hi2:])), "rows outside the range must stay untouched"


def test_adaptive_sell_hybrid_on_skewed_rows(sp, orc, monkeypatch):
    """Forced SELL candidate on an R-MAT matrix: rows longer than 256 are left out of the SELL copy and computed by the per-row /
    per-segment CTAs right after it."""
    monkeypatch.setenv("SPMVB200_FORCE_CAND", "13")
    mat = sp.synth.rmat_host_csr(17, 16)
    assert mat.MAX_ROW_NZ > 2048 and mat.NZ >= 1 << 20
    x = sp.synth.host_vector(mat.N)
    y_ref = _oracle_y(orc, mat, x)
    dm, dx, dy = sp.spMatCpyCSR(mat), sp.DeviceVector.from_host(x), sp.DeviceVector(mat.M)
    for _ in range(2):
        dy.fill_bytes(0xFF)
        sp.cudaSpMVAdaptiveCSR(dm, dx, sp.Config(), dy)
        y = dy.to_host()
        assert orc.strict_diff_csr(mat.IRP, mat.JA, mat.AS, x, y_ref, y, tau=TAU)[0] == 0
        short = np.diff(mat.IRP) <= 256
        np.testing.assert_array_equal(y[short], y_ref[short])  # the SELL part sums in the serial order
    assert dm.adaptive_choice == "sell"


def test_adaptive_hotx_hybrid_on_power_law_columns(sp, orc, monkeypatch):
    """Forced hot-x candidate (csrc/hotx.cuh) on an R-MAT matrix: ONE persistent kernel, the 16 384 hottest x entries in shared memory, column
    ids remapped; rows up to 256 entries from a SELL copy in the serial order (bit-identical), longer rows and the segments of rows longer
    than a tile by a warp each.  Repeated launches (ticket counters reset), the iterated path and the host path included."""
    monkeypatch.setenv("SPMVB200_FORCE_CAND", "14")
    if sp.capi.lib().spmvb200_get_tuning_mode() != 0:
        pytest.skip("candidates are forced through the TIMED first-use pick; this process runs with SPMVB200_TUNE=deterministic")
    mat = sp.synth.rmat_host_csr(17, 16)
    assert mat.MAX_ROW_NZ > 2048 and mat.NZ >= 1 << 20 and mat.N >= 4 * 16384
    x = sp.synth.host_vector(mat.N)
    y_ref = _oracle_y(orc, mat, x)
    dm, dx, dy = sp.spMatCpyCSR(mat), sp.DeviceVector.from_host(x), sp.DeviceVector(mat.M)
    short = np.diff(mat.IRP) <= 256
    for _ in range(3):
        dy.fill_bytes(0xFF)
        sp.cudaSpMVAdaptiveCSR(dm, dx, sp.Config(), dy)
        y = dy.to_host()
        assert orc.strict_diff_csr(mat.IRP, mat.JA, mat.AS, x, y_ref, y, tau=TAU)[0] == 0
        np.testing.assert_array_equal(y[short], y_ref[short])
    assert dm.adaptive_choice == "hotx"
    yh = np.full(mat.M, np.nan)
    sp.spmv_host(sp.CSR_ADAPTIVE, dm, x, yh)
    np.testing.assert_array_equal(yh, y)
    # the other kinds of the same handle are untouched by the remapped copy
    dy.fill_bytes(0xFF)
    sp.cudaSpMVRowsCSR(dm, dx, sp.Config(), dy)
    np.testing.assert_array_equal(dy.to_host()[np.diff(mat.IRP) <= STREAM_TILE], y_ref[np.diff(mat.IRP) <= STREAM_TILE])
    # a matrix with uniform columns does not qualify (the cache would cover < 25 % of the gathers): the pick falls back
    uni = sp.synth.host_csr(sp.synth.mixed(400000, 32, 0.1))
    du, dxu, dyu = sp.spMatCpyCSR(uni), sp.DeviceVector.from_host(sp.synth.host_vector(uni.N)), sp.DeviceVector(uni.M)
    monkeypatch.delenv("SPMVB200_FORCE_CAND")
    sp.cudaSpMVAdaptiveCSR(du, dxu, sp.Config(), dyu)
    assert du.adaptive_choice != "hotx"


def test_ell_rows_picks_sell_copy_on_padded_matrix(sp, orc, monkeypatch):
    """ELL_ROWS on mixed short/long rows (ELL rectangle >= 1.25 x nnz, >= 2^20 nnz): the handle times the column-major kernel against a
    SELL copy built from the ELL arrays; whichever runs, y is bit-identical to sgemvSerial.  With the knob set the copy is never built."""
    mat = sp.synth.host_csr(sp.synth.mixed(300000, 32, 0.05))
    assert mat.NZ >= 1 << 20 and mat.MAX_ROW_NZ * mat.M > 1.25 * mat.NZ
    ell = sp.synth.csr_to_ell_host(mat)
    x = sp.synth.host_vector(mat.N)
    y_ref = _oracle_y(orc, mat, x)
    dx, dy = sp.DeviceVector.from_host(x), sp.DeviceVector(mat.M)
    for forced_off in (False, True):
        if forced_off:
            monkeypatch.setenv("SPMVB200_ELL_NO_SELL", "1")
        dm = sp.spMatCpyELL(ell)
        assert dm.exact_choice == ""
        for _ in range(2):
            dy.fill_bytes(0xFF)
            sp.cudaSpMVRowsELL(dm, dx, sp.Config(), dy)
            np.testing.assert_array_equal(dy.to_host(), y_ref)
        assert dm.exact_choice == ("ell" if forced_off else dm.exact_choice) and dm.exact_choice in ("ell", "sell")
        # the copy also serves the other entry points of the kind: host buffers, CUDA-graph iteration is covered by test_iterate_*
        y = np.full(mat.M, np.nan)
        sp.spmv_host(sp.ELL_ROWS, dm, x, y)
        np.testing.assert_array_equal(y, y_ref)
        dm.free()


def test_exact_kind_sell_hybrid_on_skewed_rows(sp, orc, monkeypatch):
    """CSR_ROWS on an R-MAT matrix with the SELL copy forced: rows up to 256 come from the SELL copy, rows up to one tile from the
    warp-per-row kernel that adds in the serial order (csr_midrow_exact_kernel) -- both bit-identical to sgemvSerial; rows longer than
    a tile are split into segments (deterministic, within tolerance)."""
    monkeypatch.setenv("SPMVB200_FORCE_EXACT", "13")
    mat = sp.synth.rmat_host_csr(17, 16)
    lens = np.diff(mat.IRP)
    assert mat.NZ >= 1 << 20 and (lens > STREAM_TILE).any() and ((lens > 256) & (lens <= STREAM_TILE)).sum() > 10
    x = sp.synth.host_vector(mat.N)
    y_ref = _oracle_y(orc, mat, x)
    dm, dx, dy = sp.spMatCpyCSR(mat), sp.DeviceVector.from_host(x), sp.DeviceVector(mat.M)
    outs = []
    for _ in range(3):
        dy.fill_bytes(0xFF)
        sp.cudaSpMVRowsCSR(dm, dx, sp.Config(), dy)
        outs.append(dy.to_host())
    assert dm.exact_choice == "sell"
    y = outs[0]
    np.testing.assert_array_equal(y[lens <= STREAM_TILE], y_ref[lens <= STREAM_TILE])
    assert orc.strict_diff_csr(mat.IRP, mat.JA, mat.AS, x, y_ref, y, tau=TAU)[0] == 0
    for o in outs[1:]:
        np.testing.assert_array_equal(o, y)


def test_long_row_split_is_deterministic(sp):
    """Rows split across CTAs are combined in segment order by the last arriver: run-to-run identical."""
    mat = sp.synth.rmat_host_csr(14, 16)
    x = sp.synth.host_vector(mat.N)
    dx, dy, dm = sp.DeviceVector.from_host(x), sp.DeviceVector(mat.M), sp.spMatCpyCSR(mat)
    for f in (sp.cudaSpMVRowsCSR, sp.cudaSpMVAdaptiveCSR):
        outs = []
        for _ in range(5):
            dy.fill_bytes(0xFF)
            f(dm, dx, sp.Config(), dy)
            outs.append(dy.to_host())
        for o in outs[1:]:
            np.testing.assert_array_equal(o, outs[0])


def test_host_adapters_spmv_interf(sp, orc):
    """SPMV_INTERF-shaped calls with host buffers (src/include/SpMV.h:63-64), the flow of
    testSpMVImplOMP (test/SpMV_test.cu:67-101): repeated calls, result compared every time."""
    mat = sp.synth.host_csr(sp.synth.stencil27(16))
    ell = sp.synth.csr_to_ell_host(mat)
    x = sp.synth.host_vector(mat.N)
    y_ref = _oracle_y(orc, mat, x)
    for f, m in [(f, mat) for f in sp.SpmvB200CSRFuncs] + [(f, ell) for f in sp.SpmvB200ELLFuncs]:
        for _ in range(3):
            y = np.full(mat.M, np.nan)
            assert f(m, x, sp.Config(), y) == 0
            assert not orc.double_vectors_diff(y_ref, y)[0]
            assert orc.strict_diff_csr(mat.IRP, mat.JA, mat.AS, x, y_ref, y, tau=TAU)[0] == 0
            assert f.ElapsedInternal > 0
    sp.cache_drop()


@pytest.mark.parametrize("builder", [
    lambda s: s.host_csr(s.stencil27(48)),              # 110 592 rows: chunked pipeline, banded column spans
    lambda s: s.host_csr(s.banded(200000, 32, 3000)),
    lambda s: s.rmat_host_csr(17, 16),                  # rows longer than a tile inside the chunked CSR path
    lambda s: s.host_csr(s.mixed(150000, 48, 0.02)),    # uniform columns: every chunk needs all of x
])
@pytest.mark.parametrize("pinned", [False, True])
def test_pipelined_host_path(sp, orc, builder, pinned, monkeypatch):
    """spmvb200_spmv_host on matrices large enough for the chunked path: x pieces up, row chunks, y chunks down.
    pinned: y is page-locked, so the kernels store their rows straight into the caller's buffer (no device->host copy)."""
    import torch
    if pinned:
        monkeypatch.setenv("SPMVB200_HOST_DIRECT_Y", "3")  # every kind, single launches and the chunked pipeline
    mat = builder(sp.synth)
    ell = sp.synth.csr_to_ell_host(mat) if mat.MAX_ROW_NZ * mat.M < 2e7 else None
    x = sp.synth.host_vector(mat.N)
    y_ref = _oracle_y(orc, mat, x)
    short = np.diff(mat.IRP) <= STREAM_TILE
    runs = [(f, mat) for f in sp.SpmvB200CSRFuncs if not (f is sp.b200SpMVRowsXWIN and mat.MAX_ROW_NZ > 255)] + \
           ([(sp.b200SpMVRowsELL, ell)] if ell is not None else [])
    for f, m in runs:
        for rep in range(3):
            y = torch.empty(mat.M, dtype=torch.float64).pin_memory().numpy() if pinned else np.empty(mat.M)
            y.fill(np.nan)
            assert f(m, x, sp.Config(), y) == 0
            assert orc.strict_diff_csr(mat.IRP, mat.JA, mat.AS, x, y_ref, y, tau=TAU)[0] == 0, (f, rep)
            if f in (sp.b200SpMVRowsCSR, sp.b200SpMVRowsELL, sp.b200SpMVRowsSELL, sp.b200SpMVRowsXWIN):
                np.testing.assert_array_equal(y[short], y_ref[short])
            assert f.ElapsedInternal > 0
    sp.cache_drop()


def test_row_block_partition_slices(sp, orc):
    """Row-block partition: per-slice handles produce the slices of y."""
    mat = sp.synth.host_csr(sp.synth.banded(50000, 32, 3000))
    x = sp.synth.host_vector(mat.N)
    y_ref = _oracle_y(orc, mat, x)
    dx = sp.DeviceVector.from_host(x)
    bounds = [0, 12500, 25000, 37500, 50000]
    y = np.empty(mat.M)
    for a, b in zip(bounds[:-1], bounds[1:]):
        dm, dy = sp.spMatCpyCSR(mat, a, b), sp.DeviceVector(b - a)
        sp.cudaSpMVRowsCSR(dm, dx, sp.Config(), dy)
        y[a:b] = dy.to_host()
    np.testing.assert_array_equal(y, y_ref)


def test_device_generators_match_host(sp):
    s = sp.synth
    for spec in (s.lap2d(50), s.stencil27(9, 8, 7), s.banded(20000, 32, 500), s.mixed(30000, 32, 0.1)):
        h = s.host_csr(spec)
        d = s.device_csr(spec)
        irp, ja, as_ = d.download_csr()
        np.testing.assert_array_equal(irp, h.IRP)
        np.testing.assert_array_equal(ja, h.JA)
        np.testing.assert_array_equal(as_, h.AS)
    h = s.host_csr(s.banded(20000, 32, 500), 5000, 9000)
    irp, ja, as_ = s.device_csr(s.banded(20000, 32, 500), 5000, 9000).download_csr()
    np.testing.assert_array_equal(ja, h.JA)
    np.testing.assert_array_equal(as_, h.AS)
    h = s.rmat_host_csr(12, 16)
    irp, ja, as_ = s.rmat_device_csr(12, 16).download_csr()
    np.testing.assert_array_equal(irp, h.IRP)
    np.testing.assert_array_equal(ja, h.JA)
    np.testing.assert_array_equal(as_, h.AS)
    n = 1000
    dv = sp.DeviceVector(n)
    s.device_vector_fill(dv, n)
    np.testing.assert_array_equal(dv.to_host(), s.host_vector(n))


def test_full_size_cfg2_properties(sp, orc):
    """BASELINE cfg2 at full size (27-point 128^3, 55.7 M nnz), size-independent checks:
    A*1 = 27 - rowlen exactly; linearity; sampled row blocks against the oracle."""
    s = sp.synth
    spec = s.stencil27(128)
    d_csr = s.device_csr(spec)
    assert d_csr.NZ == (3 * 128 - 2) ** 3
    d_ell = d_csr.to_ell(sp.FMT_ELL_COLMAJOR)
    d_rm = d_csr.to_ell(sp.FMT_ELL_ROWMAJOR)
    M = d_csr.M
    ones = sp.DeviceVector.from_host(np.ones(M))
    dy = sp.DeviceVector(M)
    irp, _, _ = d_csr.download_csr()
    expect = 27.0 - np.diff(irp).astype(np.float64)
    for f, dm in ((sp.cudaSpMVRowsCSR, d_csr), (sp.cudaSpMVWarpPerRowCSR, d_csr), (sp.cudaSpMVAdaptiveCSR, d_csr),
                  (sp.cudaSpMVRowsELL, d_ell), (sp.cudaSpMVRowsELLNNTransposed, d_rm), (sp.cudaSpMVWarpsPerRowELLNTrasposed, d_rm)):
        dy.fill_bytes(0xFF)
        f(dm, ones, sp.Config(), dy)
        np.testing.assert_array_equal(dy.to_host(), expect, err_msg=f.__name__)
    x = s.host_vector(M)
    dx = sp.DeviceVector.from_host(x)
    sp.cudaSpMVRowsELL(d_ell, dx, sp.Config(), dy)
    y = dy.to_host()
    sp.cudaSpMVRowsCSR(d_csr, dx, sp.Config(), dy)
    np.testing.assert_array_equal(dy.to_host(), y)  # two exact kinds agree bit for bit
    for a in (0, 1_000_000, M - 50_000):  # sampled row blocks vs the oracle
        h = s.host_csr(spec, a, a + 50_000)
        np.testing.assert_array_equal(y[a:a + 50_000], orc.sgemv_serial(h.IRP, h.JA, h.AS, x))
    d2 = sp.DeviceVector.from_host(2.0 * x)  # linearity: A(2x) = 2 A x exactly (power of two)
    sp.cudaSpMVRowsELL(d_ell, d2, sp.Config(), dy)
    np.testing.assert_array_equal(dy.to_host(), 2.0 * y)
