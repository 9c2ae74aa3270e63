"""CPU-only checks of the product boundary: the C-ABI library builds/loads, exports every symbol
include/spmv_b200.h declares, fails loudly without a GPU, and the host generators are deterministic.
No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden


@pytest.fixture(scope="module")
def sp():
    import spmv_openmp_cuda_b200 as sp
    sp.capi.lib()
    return sp


def test_header_symbols_all_exported(sp):
    hdr = open(os.path.join(ROOT, "include", "spmv_b200.h")).read()
    hdr = hdr.split("reference-typed adapters")[0]
    declared = set(re.findall(r"\b(spmvb200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 35
    L = C.CDLL(sp.capi.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(sp.capi.EXPORTS), declared ^ set(sp.capi.EXPORTS)


def test_version_and_kind_names(sp):
    L = sp.capi.lib()
    assert L.spmvb200_version() == 200
    names = [L.spmvb200_kind_name(k).decode() for k in range(6)]
    assert names[0] == "CUDA_CSR_ROWS" and names[1] == "CUDA_CSR_ROWS_WARP" and names[2] == "CUDA_ELL_ROWS"
    assert names[4] == "CUDA_ELL_ROWS_WARP_NN_TRANSPOSED"  # src/include/SpMV.h:41


def test_tables_mirror_reference(sp):
    # src/include/SpMV.h:130-142
    assert [f.__name__ for f in sp.SpmvCUDA_CSRFuncs[:2]] == ["cudaSpMVRowsCSR", "cudaSpMVWarpPerRowCSR"]
    assert [f.__name__ for f in sp.SpmvCUDA_ELLFuncs] == ["cudaSpMVRowsELL", "cudaSpMVRowsELLNNTransposed",
                                                          "cudaSpMVWarpsPerRowELLNTrasposed"]
    assert sp.SpmvCUDA_CSRFuncs_WarpPerRowIdx == 1 and sp.SpmvCUDA_ELLFuncs_WarpPerRowIdx == 2


def test_no_cpu_fallback(sp):
    """Without a device the product path must raise, never compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    g = load_golden("lap2d_8")
    mat = sp.Spmat.csr(int(g["N"]), g["irp"], g["ja"], g["as_"])
    with pytest.raises(sp.SpmvB200Error):
        sp.spMatCpyCSR(mat)
    with pytest.raises(sp.SpmvB200Error):
        sp.b200SpMVRowsCSR(mat, g["x"], sp.Config(), np.zeros(mat.M))
    with pytest.raises(sp.SpmvB200Error):
        sp.synth.device_csr(sp.synth.lap2d(8))
    # host-buffer helpers: no device, no page-locking -- loud failure, nothing registered
    L = sp.capi.lib()
    buf = np.zeros(1 << 18)
    assert L.spmvb200_host_alloc(1 << 20) is None and L.spmvb200_last_error()
    assert L.spmvb200_host_register(sp.capi.ptr(buf), buf.nbytes) != 0
    assert L.spmvb200_host_registered(sp.capi.ptr(buf)) == 0
    assert L.spmvb200_host_unregister(None) == 0 and L.spmvb200_host_free(None) == 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "spmv_openmp_cuda_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_host_generators_match_reference_parser(sp):
    """The 5-point generator reproduces, entry for entry, what the reference's Matrix Market reader
    built for the same matrix (golden fixture written by tests/golden/make_golden.py)."""
    for n, name in ((8, "lap2d_8"), (33, "lap2d_33")):
        g = load_golden(name)
        m = sp.synth.host_csr(sp.synth.lap2d(n))
        np.testing.assert_array_equal(m.IRP, g["irp"])
        np.testing.assert_array_equal(m.JA, g["ja"])
        np.testing.assert_array_equal(m.AS, g["as_"])
        e = sp.synth.csr_to_ell_host(m)
        np.testing.assert_array_equal(e.JA, g["ell_ja"])
        np.testing.assert_array_equal(e.AS, g["ell_as"])


def test_generators_row_ranges_and_shapes(sp):
    s = sp.synth
    full = s.host_csr(s.banded(4096, 32, 300))
    assert full.NZ == 4096 * 32
    ja = full.JA.reshape(4096, 32).astype(np.int64)
    assert np.all(np.diff(ja, axis=1) > 0)  # sorted, distinct
    rows = np.arange(4096)[:, None]
    assert np.all(np.abs(ja - rows) <= 300) and ja.min() >= 0 and ja.max() < 4096
    part = s.host_csr(s.banded(4096, 32, 300), 1000, 1500)
    np.testing.assert_array_equal(part.JA, full.JA[1000 * 32:1500 * 32])
    np.testing.assert_array_equal(part.AS, full.AS[1000 * 32:1500 * 32])
    st = s.host_csr(s.stencil27(7, 6, 5))
    assert st.NZ == (3 * 7 - 2) * (3 * 6 - 2) * (3 * 5 - 2) and st.MAX_ROW_NZ == 27
    mx = s.host_csr(s.mixed(20000, 48, 0.05))
    rl = np.diff(mx.IRP)
    assert set(rl.tolist()) == {4, 48} and 0.03 < (rl == 48).mean() < 0.07
    for r in (0, 7, 19999):
        c = mx.JA[int(mx.IRP[r]):int(mx.IRP[r + 1])].astype(np.int64)
        assert np.all(np.diff(c) > 0)
    rm = s.rmat_host_csr(9, 8)
    assert rm.M == 512 and rm.NZ <= 8 * 512 and rm.NZ > 2000
    for r in range(0, 512, 37):
        c = rm.JA[int(rm.IRP[r]):int(rm.IRP[r + 1])].astype(np.int64)
        assert np.all(np.diff(c) > 0)
    x1, x2 = s.host_vector(100), s.host_vector(40, begin=60)
    np.testing.assert_array_equal(x1[60:], x2)
    assert np.all(np.abs(x1) < 3e-5) and np.all(np.isfinite(x1))


def test_binary_matrix_cache_roundtrip(tmp_path):
    """Spmat.save / Spmat.load: CSR and ELL round-trip bit for bit; garbage is refused."""
    import spmv_openmp_cuda_b200 as sp
    rng = np.random.default_rng(3)
    lens = rng.integers(0, 9, 50)
    irp = np.zeros(51, dtype=np.uint64)
    irp[1:] = np.cumsum(lens)
    ja = rng.integers(0, 70, int(irp[-1])).astype(np.uint64)
    csr = sp.Spmat.csr(70, irp, ja, rng.uniform(-1, 1, int(irp[-1])))
    ell = sp.Spmat.ell(4, 6, 3, np.arange(12, dtype=np.uint64) % 6, rng.uniform(-1, 1, 12), RL=np.array([3, 1, 0, 2], dtype=np.uint64))
    for i, m in enumerate((csr, ell)):
        path = str(tmp_path / ("m%d.bin" % i))
        m.save(path)
        r = sp.Spmat.load(path)
        assert (r.M, r.N, r.NZ, r.MAX_ROW_NZ) == (m.M, m.N, m.NZ, m.MAX_ROW_NZ)
        for a, b in ((r.IRP, m.IRP), (r.JA, m.JA), (r.AS, m.AS), (r.RL, m.RL)):
            assert (a is None) == (b is None)
            if a is not None:
                np.testing.assert_array_equal(a, b)
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"%%MatrixMarket matrix coordinate real general\n")
    with pytest.raises(ValueError):
        sp.Spmat.load(str(bad))


def test_bench_reference_arm_contract_line():
    """`bench.py --impl reference` (the reference's own OpenMP kernels on the host cores) prints exactly one JSON line with the
    contract's keys; `bench.py` without a GPU refuses to run (no CPU fallback on the product arm)."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["dtype"] == "f64" and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    import torch
    if not torch.cuda.is_available():
        ours = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=600)
        assert ours.returncode != 0 and "no CPU fallback" in (ours.stderr + ours.stdout)


# ---------------------------------------------------------------------------------------- round 2: comparators, CLI
def test_strict_comparator_in_the_c_abi_matches_the_oracle(oracle_mod):
    """spmvb200_compare_strict_csr / _compare_abs (host arrays, no GPU needed) against the oracle's comparators: same verdicts on
    exact, slightly perturbed, badly perturbed and NaN outputs."""
    import ctypes as C

    from spmv_openmp_cuda_b200 import capi, synth
    lib = capi.lib()
    mat = synth.host_csr(synth.mixed(3000, 24, 0.2))
    x = synth.host_vector(mat.N)
    y_ref = oracle_mod.sgemv_serial(mat.IRP, mat.JA, mat.AS, x)
    cases = {"exact": y_ref.copy(), "ulp": y_ref * (1 + 2e-16), "off": y_ref.copy(), "nan": y_ref.copy(), "inf": y_ref.copy()}
    cases["off"][17] += 1e-9
    cases["nan"][5] = np.nan
    cases["inf"][9] = np.inf
    for name, y in cases.items():
        bad, worst = C.c_uint64(99), C.c_double(-1)
        rc = lib.spmvb200_compare_strict_csr(mat.M, capi.ptr(mat.IRP), capi.ptr(mat.JA), capi.ptr(mat.AS), capi.ptr(x), capi.ptr(y_ref), capi.ptr(y), 1e-12,
                                             C.byref(bad), C.byref(worst))
        assert rc == 0
        want_bad, want_worst = oracle_mod.strict_diff_csr(mat.IRP, mat.JA, mat.AS, x, y_ref, y, tau=1e-12)
        assert (bad.value > 0) == (want_bad > 0) == (name in ("off", "nan", "inf")), (name, bad.value, want_bad)
        failed, dmax = C.c_int(-1), C.c_double(-1)
        assert lib.spmvb200_compare_abs(mat.M, capi.ptr(y_ref), capi.ptr(y), 7e-4, C.byref(failed), C.byref(dmax)) == 0
        assert failed.value == (1 if name in ("nan", "inf") else 0), name  # 1e-9 passes the reference's absolute 7e-4; NaN must not
    # an all-zero row with a zero output is fine, with a non-zero output it is not
    irp = np.array([0, 0, 1], dtype=np.uint64)
    ja, as_, xx = np.array([0], dtype=np.uint64), np.array([2.0]), np.array([3.0])
    for y, want in ((np.array([0.0, 6.0]), 0), (np.array([1e-30, 6.0]), 1)):
        bad = C.c_uint64(99)
        assert lib.spmvb200_compare_strict_csr(2, capi.ptr(irp), capi.ptr(ja), capi.ptr(as_), capi.ptr(xx), capi.ptr(np.array([0.0, 6.0])), capi.ptr(y), 1e-12,
                                               C.byref(bad), None) == 0
        assert bad.value == want


def _write_mtx(path, m):
    rows = np.repeat(np.arange(m.M), np.diff(m.IRP).astype(np.int64))
    with open(path, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (m.M, m.N, m.NZ))
        for r, c, v in zip(rows, m.JA, m.AS):
            f.write("%d %d %.17g\n" % (r + 1, c + 1, v))


def test_cli_driver_modes_on_cpu(tmp_path):
    """tests/integration/b200_main: the reference's command line (src/main.cu:69-139).  Without a GPU the OMP modes run (they are
    the unmodified reference functions), the B200 modes fail LOUDLY (no CPU fallback), an unknown mode prints the usage."""
    import subprocess

    from spmv_openmp_cuda_b200 import synth
    exe = os.path.join(ROOT, "tests", "integration", "_build", "b200_main")
    if not os.path.exists(exe):
        pytest.skip("b200_main not built (needs /root/reference at build time)")
    p = str(tmp_path / "m.mtx")
    _write_mtx(p, synth.host_csr(synth.lap2d(40)))
    env = dict(os.environ, OMP_SCHEDULE="nonmonotonic:static", CUDA_VISIBLE_DEVICES="")
    for mode in ("CSR_ROWS", "CSR_ROWS_GROUPS", "ELL_ROWS_GROUPS"):
        out = subprocess.run([exe, p, "RNDVECT", mode, "--check"], capture_output=True, text=True, env=env, timeout=120)
        assert out.returncode == 0, out.stderr[-500:]
        assert "cmode:" in out.stdout and "elapsedInternal" in out.stdout and "doubleVectorsDiff ok" in out.stdout
    for mode in ("B200_CSR_ROWS", "CUDA_CSR_ROWS_WARP", "B200_ELL_ROWS"):
        out = subprocess.run([exe, p, "RNDVECT", mode], capture_output=True, text=True, env=env, timeout=120)
        assert out.returncode != 0 and "no CUDA device" in out.stderr, (mode, out.stderr[-500:])
    out = subprocess.run([exe, p, "RNDVECT", "NOT_A_MODE"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode != 0 and "INVALID COMPUTE_MODE" in out.stderr and "B200_CSR_ADAPTIVE" in out.stderr


def test_reference_log_parser_reads_the_harness_log():
    """scripts/parseLog.py of the reference, UNMODIFIED, run on a log the C harness (tests/integration/b200_harness) wrote on the
    B200 box (committed fixture): it must yield one CSV row per B200 implementation with the matrix sizes and times filled in."""
    import subprocess
    import sys
    parser = "/root/reference/scripts/parseLog.py"
    log = os.path.join(ROOT, "tests", "golden", "b200_harness_lap2d_150.log")
    if not os.path.exists(parser) or not os.path.exists(log):
        pytest.skip("needs the reference tree and the committed harness log")
    out = subprocess.run([sys.executable, parser, log], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr[-1000:]
    rows = [ln for ln in out.stdout.splitlines() if "B200" in ln]
    assert len(rows) == 7, out.stdout
    hdr = out.stdout.splitlines()[0].split(",")
    for ln in rows:
        f = [v.strip() for v in ln.split(",")]
        rec = dict(zip(hdr, f))
        assert rec["matRows"] == "22500" and rec["matCols"] == "22500" and int(rec["NNZ"]) == 5 * 150 * 150 - 4 * 150
        assert float(rec["timeAvg"]) > 0 and rec["sampleSize"] in ("5", "25")  # AVG_TIMES_ITERATION of the harness build
