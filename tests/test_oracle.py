"""The oracle (oracle/spmv_oracle.c) pinned against the reference:
 (1) the committed golden fixtures, produced by the unmodified reference (tests/golden/make_golden.py);
 (2) where oracle/_ref/libspmv_ref.so exists (this container; it also travels to the GPU box),
     the reference itself on fresh seeded inputs.
CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden

NAMES = golden_names()


def test_golden_present():
    assert len(NAMES) >= 8


@pytest.mark.parametrize("name", NAMES)
def test_mm_to_csr_matches_reference_parser(oracle_mod, name):
    g = load_golden(name)
    M, N, row, col, val = oracle_mod.mm_to_coo(str(g["mtx"]))
    assert (M, N, len(row)) == (int(g["M"]), int(g["N"]), int(g["NZ"]))
    irp, ja, as_, rl = oracle_mod.coo_to_csr(M, row, col, val)
    np.testing.assert_array_equal(irp, g["irp"])
    np.testing.assert_array_equal(ja, g["ja"])
    np.testing.assert_array_equal(as_, g["as_"])
    np.testing.assert_array_equal(rl, g["rl"])


@pytest.mark.parametrize("name", NAMES)
def test_mm_to_ell_and_transpose_match_reference(oracle_mod, name):
    g = load_golden(name)
    M, N, row, col, val = oracle_mod.mm_to_coo(str(g["mtx"]))
    K, ja, as_, rl = oracle_mod.coo_to_ell(M, row, col, val)
    assert K == int(g["K"])
    np.testing.assert_array_equal(ja, g["ell_ja"])
    np.testing.assert_array_equal(as_, g["ell_as"])
    np.testing.assert_array_equal(rl, g["rl"])
    ja_t, as_t = oracle_mod.ell_transpose(M, K, ja, as_)
    np.testing.assert_array_equal(ja_t, g["ell_ja_t"])
    np.testing.assert_array_equal(as_t, g["ell_as_t"])
    # the reference swaps M <-> MAX_ROW_NZ on the transposed struct (sparseUtils.c:168-171)
    np.testing.assert_array_equal(g["ell_t_dims"], np.array([K, M, M], dtype=np.uint64))


@pytest.mark.parametrize("name", NAMES)
def test_sgemv_serial_bit_exact_vs_reference(oracle_mod, name):
    g = load_golden(name)
    y = oracle_mod.sgemv_serial(g["irp"], g["ja"], g["as_"], g["x"])
    np.testing.assert_array_equal(y, g["y_sgemvSerial"])  # bit exact
    yb = oracle_mod.spmv_rows_blocks_csr(g["irp"], g["ja"], g["as_"], g["x"], grid_rows=min(8, max(1, len(y))))
    np.testing.assert_array_equal(yb, g["y_sgemvSerial"])
    yr = oracle_mod.spmv_rows_basic_csr(g["irp"], g["ja"], g["as_"], g["x"])
    np.testing.assert_array_equal(yr, g["y_sgemvSerial"])


@pytest.mark.parametrize("name", NAMES)
def test_ell_paths_vs_reference(oracle_mod, name):
    g = load_golden(name)
    M, K = int(g["M"]), int(g["K"])
    y_rl = oracle_mod.spmv_rows_ell(M, K, g["ell_ja"], g["ell_as"], g["x"], rl=g["rl"])
    y_full = oracle_mod.spmv_rows_ell(M, K, g["ell_ja"], g["ell_as"], g["x"])
    y_t = oracle_mod.spmv_ell_colmajor(M, K, M, g["ell_ja_t"], g["ell_as_t"], g["x"])
    # the reference's ELL OMP kernels were built with `omp simd reduction` (config.h:92-94): same
    # terms, possibly another order => compare with the strict relative bound, and the serial
    # oracle bit-for-bit (row-major ELL with row lengths visits exactly the CSR terms in order).
    np.testing.assert_array_equal(y_rl, g["y_sgemvSerial"])
    np.testing.assert_array_equal(y_full, y_t)
    for key in ("y_spmvRowsBasicELL", "y_spmvRowsBlocksELL", "y_spmvTilesELL"):
        bad, worst = oracle_mod.strict_diff_csr(g["irp"], g["ja"], g["as_"], g["x"], g[key], y_rl, tau=1e-12)
        assert bad == 0, (key, worst)
    bad, worst = oracle_mod.strict_diff_csr(g["irp"], g["ja"], g["as_"], g["x"], g["y_sgemvSerial"], y_full, tau=1e-12)
    assert bad == 0, worst


@pytest.mark.parametrize("name", NAMES)
def test_reference_omp_variants_within_stated_tolerance(oracle_mod, name):
    """All reference CPU implementations agree with the oracle: reference check (abs 7e-4) and
    this repo's strict check (tau = 1e-12 relative to sum|a||x|)."""
    g = load_golden(name)
    yref = oracle_mod.sgemv_serial(g["irp"], g["ja"], g["as_"], g["x"])
    for key in [k for k in g if k.startswith("y_")]:
        failed, dmax = oracle_mod.double_vectors_diff(yref, g[key])
        assert not failed, (key, dmax)
        bad, worst = oracle_mod.strict_diff_csr(g["irp"], g["ja"], g["as_"], g["x"], yref, g[key], tau=1e-12)
        assert bad == 0, (key, worst)


def test_comparators(oracle_mod):
    a = np.array([0.0, 1.0, 2.0])
    assert oracle_mod.double_vectors_diff(a, a + 6e-4)[0] is False
    failed, dmax = oracle_mod.double_vectors_diff(a, a + np.array([0, 8e-4, 0]))
    assert failed and abs(dmax + 8e-4) < 1e-12
    # NaN-blind like the reference (utils.c:368-370) ...
    assert oracle_mod.double_vectors_diff(a, np.array([0.0, np.nan, 2.0]))[0] is False
    # ... which is why the strict comparator exists
    irp = np.array([0, 1, 2, 3], dtype=np.uint64)
    ja = np.array([0, 1, 2], dtype=np.uint64)
    as_ = np.ones(3)
    x = np.ones(3)
    y = np.ones(3)
    assert oracle_mod.strict_diff_csr(irp, ja, as_, x, y, y)[0] == 0
    assert oracle_mod.strict_diff_csr(irp, ja, as_, x, y, np.array([1.0, np.nan, 1.0]))[0] == 1
    assert oracle_mod.strict_diff_csr(irp, ja, as_, x, y, y * (1 + 1e-10))[0] == 3
    avg, var = oracle_mod.stats_avg_var([1.0, 2.0, 3.0, 4.0])
    assert avg == 2.5 and abs(var - 1.25) < 1e-15


def test_unsorted_coo_rejected(oracle_mod):
    row = np.array([0, 0], dtype=np.uint64)
    col = np.array([3, 1], dtype=np.uint64)
    with pytest.raises(ValueError):
        oracle_mod.coo_to_csr(1, row, col, np.ones(2))
    irp, ja, _, _ = oracle_mod.coo_to_csr(1, row, col, np.ones(2), check_sorted=False)
    assert list(ja) == [3, 1] and list(irp) == [0, 2]


# ------------------------------------------------------------------ live reference (when built)
def _rand_csr(rng, M, N, mean_len):
    lens = rng.poisson(mean_len, M).clip(0, N)
    irp = np.zeros(M + 1, dtype=np.uint64)
    irp[1:] = np.cumsum(lens)
    ja = np.concatenate([np.sort(rng.choice(N, k, replace=False)) for k in lens] + [np.zeros(0, int)]).astype(np.uint64)
    as_ = rng.uniform(-1, 1, int(irp[-1]))
    return irp, ja, as_, lens.astype(np.uint64)


@pytest.mark.parametrize("seed,M,N,mean_len", [(1, 500, 400, 7), (2, 2000, 2000, 1.5), (3, 64, 5000, 300)])
def test_live_reference_matches_oracle(oracle_mod, seed, M, N, mean_len):
    if not oracle_mod.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(seed)
    irp, ja, as_, rl = _rand_csr(rng, M, N, mean_len)
    x = rng.uniform(-1, 1, N)
    mat = oracle_mod.ref_spmat(M, N, int(irp[-1]), ja, as_, irp=irp, rl=rl)
    cfg = oracle_mod.ref_config(grid_rows=8, grid_cols=8, chunks=0)
    y_ref = oracle_mod.ref_call("sgemvSerial", mat, x, cfg, M)
    np.testing.assert_array_equal(oracle_mod.sgemv_serial(irp, ja, as_, x), y_ref)
    for fn in ("spmvRowsBasicCSR", "spmvRowsBlocksCSR", "spmvTilesCSR", "spmvTilesAllocdCSR"):
        y = oracle_mod.ref_call(fn, mat, x, cfg, M)
        bad, worst = oracle_mod.strict_diff_csr(irp, ja, as_, x, y_ref, y, tau=1e-12)
        assert bad == 0, (fn, worst)
    # the reference comparator itself
    dm = C.c_double(0)
    y2 = y_ref.copy()
    y2[M // 2] += 1e-3
    assert oracle_mod.ref().doubleVectorsDiff(y_ref, y2, M, C.byref(dm)) != 0
    assert oracle_mod.double_vectors_diff(y_ref, y2)[0] is True
    assert abs(dm.value - oracle_mod.double_vectors_diff(y_ref, y2)[1]) == 0.0


def test_cblas_cross_check_of_the_oracle(oracle_mod):
    """The reference pins sgemvSerial with a dense cblas_dgemv on small cases (test/SpMV_CBLAS.c:32-57, test/SpMV_test.cu:221-236).
    Same check here, with the reference's own prebuilt reference-BLAS archives (oracle/_ref/cblas_check, built by
    `make -C oracle cblas`): the restatement, the live reference and CBLAS agree on every golden fixture and on seeded random
    matrices -- to 1e-13 relative to the row's magnitude (dgemv accumulates in another order)."""
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "cblas_check")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/cblas_check not built (needs the reference's CBLAS archives)")

    def dgemv(dense, x):
        M, N = dense.shape
        blob = np.array([M, N], dtype=np.uint64).tobytes() + np.ascontiguousarray(dense).tobytes() + np.ascontiguousarray(x).tobytes()
        out = subprocess.run([exe], input=blob, capture_output=True, timeout=60)
        assert out.returncode == 0, out.stderr[-300:]
        return np.frombuffer(out.stdout, dtype=np.float64)

    cases = []
    for name in golden_names():
        g = load_golden(name)
        cases.append((name, int(g["M"]), int(g["N"]), g["irp"], g["ja"], g["as_"], g["x"]))
    rng = np.random.default_rng(5)
    for k in range(3):
        M, N = int(rng.integers(20, 90)), int(rng.integers(20, 90))
        lens = rng.integers(0, min(N, 12), M)
        irp = np.zeros(M + 1, dtype=np.uint64)
        irp[1:] = np.cumsum(lens)
        ja = np.concatenate([np.sort(rng.choice(N, int(n), replace=False)) for n in lens] + [np.zeros(0, int)]).astype(np.uint64)
        cases.append(("rand%d" % k, M, N, irp, ja, rng.uniform(-1, 1, int(irp[-1])), rng.uniform(-1, 1, N)))
    for name, M, N, irp, ja, as_, x in cases:
        if M * N > 4_000_000:
            continue
        dense = np.zeros((M, N))
        absd = np.zeros((M, N))
        rows = np.repeat(np.arange(M), np.diff(irp).astype(np.int64))
        np.add.at(dense, (rows, ja.astype(np.int64)), as_)  # duplicates (if any) accumulate, as CSRToDense would overwrite: fixtures have none
        np.add.at(absd, (rows, ja.astype(np.int64)), np.abs(as_))
        y_blas = dgemv(dense, x)
        y_or = oracle_mod.sgemv_serial(irp, ja, as_, x)
        scale = absd @ np.abs(x) + 1e-300
        assert np.all(np.abs(y_blas - y_or) <= 1e-13 * scale), name
        if oracle_mod.ref_available():
            rm = oracle_mod.ref_spmat(M, N, int(irp[-1]), ja, as_, irp=irp, rl=np.diff(irp).astype(np.uint64))
            y_ref = oracle_mod.ref_call("sgemvSerial", rm, x, oracle_mod.ref_config(), M)
            np.testing.assert_array_equal(y_ref, y_or)
            assert np.all(np.abs(y_blas - y_ref) <= 1e-13 * scale), name
