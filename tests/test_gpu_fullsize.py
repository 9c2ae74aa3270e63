"""Parity at the BASELINE.json FULL sizes (cfg1, cfg3, cfg4, cfg5; cfg2 is in test_gpu_parity.py).

Full size changes the code path -- first-use picks build x-window / SELL copies with their default geometry, the 16-bit offset
range check runs over 2^25 rows, 2^30 non-zeros sit close to the 32-bit offset limit -- so the scaled-down cases do not cover it.
Matrices are generated on the device; the oracle (sgemvSerial, src/SpMV_CSR_OMP.c:229-250) sees either the whole matrix
regenerated on the host (cfg1, cfg3) or sampled row blocks of it (cfg4, cfg5: any row range of a synthetic matrix can be
regenerated on its own).  Tolerances as in test_gpu_parity.py: the exact kinds bit for bit WHATEVER the first-use pick was,
the tree-reduced kinds within 1e-12 * sum|a_ij x_j|.
"""
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


def _check_blocks(sp, orc, spec, blocks, x, y, exact, name):
    for a, b in blocks:
        h = sp.synth.host_csr(spec, a, b)
        want = orc.sgemv_serial(h.IRP, h.JA, h.AS, x)
        got = y[a:b]
        assert np.all(np.isfinite(got)), name
        if exact:
            short = np.diff(h.IRP) <= STREAM_TILE
            np.testing.assert_array_equal(got[short], want[short], err_msg=name)
        bad, worst = orc.strict_diff_csr(h.IRP, h.JA, h.AS, x, want, got, tau=TAU)
        assert bad == 0, (name, a, b, bad, worst)


def test_full_size_cfg1_all_kinds(sp, orc):
    """cfg1: 5-point Laplacian 1024^2 (1 048 576 rows, 5 238 784 nnz): whole-matrix comparison for every kind."""
    s = sp.synth
    spec = s.lap2d(1024)
    mat = s.host_csr(spec)
    assert mat.NZ == 5 * 1024 * 1024 - 4 * 1024
    x = s.host_vector(mat.N)
    y_ref = orc.sgemv_serial(mat.IRP, mat.JA, mat.AS, x)
    d_csr = s.device_csr(spec)
    d_ell, d_rm = d_csr.to_ell(sp.FMT_ELL_COLMAJOR), d_csr.to_ell(sp.FMT_ELL_ROWMAJOR)
    d_sell, d_xw = d_csr.to_sell(), d_csr.to_xwin()
    dx, dy = sp.DeviceVector.from_host(x), sp.DeviceVector(mat.M)
    for f, dm, exact in ((sp.cudaSpMVRowsCSR, d_csr, True), (sp.cudaSpMVWarpPerRowCSR, d_csr, False), (sp.cudaSpMVAdaptiveCSR, d_csr, False),
                         (sp.cudaSpMVRowsELL, d_ell, True), (sp.cudaSpMVRowsELLNNTransposed, d_rm, False),
                         (sp.cudaSpMVWarpsPerRowELLNTrasposed, d_rm, False), (sp.cudaSpMVRowsSELL, d_sell, True), (sp.cudaSpMVRowsXWIN, d_xw, True)):
        dy.fill_bytes(0xFF)
        f(dm, dx, sp.Config(), dy)
        y = dy.to_host()
        if exact:
            np.testing.assert_array_equal(y, y_ref, err_msg=f.__name__)
        bad, worst = orc.strict_diff_csr(mat.IRP, mat.JA, mat.AS, x, y_ref, y, tau=TAU)
        assert bad == 0, (f.__name__, bad, worst)
    assert d_ell.index_bits == 16  # 5-point stencil: columns within 2^16 of the row


def test_full_size_cfg3_rmat_scale22(sp, orc):
    """cfg3: R-MAT scale 22, edge factor 16 (4 194 304 rows, ~6.4e7 nnz after merging duplicates, longest row ~1.6e5): the
    device-built matrix equals the host-built one, every CSR kind against the oracle on the WHOLE matrix; the exact kind is
    bit-identical on every row of at most 2048 non-zeros whatever it picked (stream / SELL hybrid)."""
    s = sp.synth
    mat = s.rmat_host_csr(22, 16)
    d_csr = s.rmat_device_csr(22, 16)
    assert (d_csr.M, d_csr.NZ) == (mat.M, mat.NZ)
    x = s.host_vector(mat.N)
    y_ref = orc.sgemv_serial(mat.IRP, mat.JA, mat.AS, x)
    lens = np.diff(mat.IRP)
    assert lens.max() > 50_000  # the heavy skew the config asks for
    short = lens <= STREAM_TILE
    dx, dy = sp.DeviceVector.from_host(x), sp.DeviceVector(mat.M)
    d_sell = d_csr.to_sell()
    for f, dm, exact in ((sp.cudaSpMVRowsCSR, d_csr, True), (sp.cudaSpMVWarpPerRowCSR, d_csr, False), (sp.cudaSpMVAdaptiveCSR, d_csr, False),
                         (sp.cudaSpMVRowsSELL, d_sell, True)):
        dy.fill_bytes(0xFF)
        f(dm, dx, sp.Config(), dy)
        y = dy.to_host()
        assert np.all(np.isfinite(y)), f.__name__
        if exact:
            np.testing.assert_array_equal(y[short], y_ref[short], err_msg=f.__name__)
            y2 = y.copy()
            dy.fill_bytes(0xFF)
            f(dm, dx, sp.Config(), dy)
            np.testing.assert_array_equal(dy.to_host(), y2, err_msg=f.__name__ + " run-to-run")  # long rows: deterministic
        bad, worst = orc.strict_diff_csr(mat.IRP, mat.JA, mat.AS, x, y_ref, y, tau=TAU)
        assert bad == 0, (f.__name__, bad, worst)


def test_full_size_cfg4_banded_2p30(sp, orc):
    """cfg4: random banded 2^25 rows x 32 nnz/row = 2^30 nnz, w = 2^15 (the bench workload): sampled row blocks at the start, the
    end, the middle and around the 2^31-byte / 2^32-byte offsets of the value array, for the three CSR kinds."""
    s = sp.synth
    M = 1 << 25
    spec = s.banded(M, 32, 1 << 15)
    d_csr = s.device_csr(spec)
    assert d_csr.NZ == 1 << 30
    x = s.host_vector(M)
    dx, dy = sp.DeviceVector.from_host(x), sp.DeviceVector(M)
    B = 20_000
    blocks = [(0, B), (M - B, M), (M // 2 - B // 2, M // 2 + B // 2), ((1 << 23) - B // 2, (1 << 23) + B // 2), ((1 << 24) - B // 2, (1 << 24) + B // 2),
              (12_345_678, 12_345_678 + B)]
    ys = {}
    for f, exact in ((sp.cudaSpMVRowsCSR, True), (sp.cudaSpMVWarpPerRowCSR, False), (sp.cudaSpMVAdaptiveCSR, False)):
        dy.fill_bytes(0xFF)
        f(d_csr, dx, sp.Config(), dy)
        y = dy.to_host()
        assert np.all(np.isfinite(y)), f.__name__
        _check_blocks(sp, orc, spec, blocks, x, y, exact, f.__name__)
        ys[f.__name__] = y
    assert d_csr.exact_choice in ("xwindow", "stream", "sell")
    # the tolerance kinds agree with the exact kind everywhere (not only in the sampled blocks) to 1e-12 of the row's magnitude:
    # |a_ij| < 1, |x_j| < 3e-5, 32 entries per row => sum|a x| < 1e-3
    for k in ("cudaSpMVWarpPerRowCSR", "cudaSpMVAdaptiveCSR"):
        assert np.max(np.abs(ys[k] - ys["cudaSpMVRowsCSR"])) <= 1e-15
    # linearity: A(2x) = 2 A x exactly (power of two), on the exact kind
    d2 = sp.DeviceVector.from_host(2.0 * x)
    sp.cudaSpMVRowsCSR(d_csr, d2, sp.Config(), dy)
    np.testing.assert_array_equal(dy.to_host(), 2.0 * ys["cudaSpMVRowsCSR"])


@pytest.mark.parametrize("p_long,rho", [(0.6364, 1.5), (0.0455, 8.0)])
def test_full_size_cfg5_mixed_rows_k48(sp, orc, p_long, rho):
    """cfg5: 2^23 rows, short rows of 4, a fraction p of rows of K_max = 48, uniform random columns; two padding ratios
    rho = M*K_max/nnz.  ELL (with and without the SELL copy it may build) and CSR kinds on sampled row blocks."""
    s = sp.synth
    M = 1 << 23
    spec = s.mixed(M, 48, p_long)
    d_csr = s.device_csr(spec)
    assert abs(M * 48 / d_csr.NZ - rho) < 0.05 * rho
    d_ell = d_csr.to_ell(sp.FMT_ELL_COLMAJOR)
    d_sell = d_csr.to_sell()
    x = s.host_vector(M)
    dx, dy = sp.DeviceVector.from_host(x), sp.DeviceVector(M)
    B = 20_000
    blocks = [(0, B), (M - B, M), (M // 2, M // 2 + B), (3_333_333, 3_333_333 + B)]
    y_exact = None
    for f, dm, exact in ((sp.cudaSpMVRowsCSR, d_csr, True), (sp.cudaSpMVRowsELL, d_ell, True), (sp.cudaSpMVRowsSELL, d_sell, True),
                         (sp.cudaSpMVWarpPerRowCSR, d_csr, False), (sp.cudaSpMVAdaptiveCSR, d_csr, False)):
        dy.fill_bytes(0xFF)
        f(dm, dx, sp.Config(), dy)
        y = dy.to_host()
        _check_blocks(sp, orc, spec, blocks, x, y, exact, f.__name__)
        if exact:
            if y_exact is None:
                y_exact = y
            np.testing.assert_array_equal(y, y_exact, err_msg=f.__name__)  # the exact kinds agree on EVERY row
    assert d_ell.exact_choice in ("ell", "sell")
