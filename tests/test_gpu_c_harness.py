"""The C drop-in, end to end on the GPU: tests/integration/b200_harness (built in the container against the
REAL reference headers and oracle/_ref/libspmv_ref.so) parses a Matrix Market file with the reference's reader,
computes the oracle with the reference's sgemvSerial, runs the b200SpMV* SPMV_INTERF adapters of
include/spmv_b200.h and compares with the reference's doubleVectorsDiff after every repetition."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
HARNESS = os.path.join(ROOT, "tests", "integration", "_build", "b200_harness")


def _write_mtx(path, M, N, rows, cols, vals):
    with open(path, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (M, N, len(rows)))
        for r, c, v in zip(rows, cols, vals):
            f.write("%d %d %.17g\n" % (r + 1, c + 1, v))


@pytest.mark.parametrize("case", ["lap2d_150", "skewed"])
def test_reference_side_harness(tmp_path, case):
    if not os.path.exists(HARNESS):
        pytest.skip("harness not built (needs /root/reference at build time)")
    import spmv_openmp_cuda_b200 as sp
    sp.capi.require_device()
    if case == "lap2d_150":
        m = sp.synth.host_csr(sp.synth.lap2d(150))
    else:
        m = sp.synth.rmat_host_csr(12, 12)
    rows = np.repeat(np.arange(m.M), np.diff(m.IRP).astype(np.int64))
    p = str(tmp_path / (case + ".mtx"))
    _write_mtx(p, m.M, m.N, rows, m.JA, m.AS)
    env = dict(os.environ, OMP_SCHEDULE="nonmonotonic:static")
    out = subprocess.run([HARNESS, p], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "B200_HARNESS_OK" in out.stdout, out.stdout[-1500:] + out.stderr[-1500:]
    # the reference's log format (test/SpMV_test.cu:93-96) so scripts/parseLog.py keeps working
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("threadNum:")]
    assert len(lines) == 7 and all("timeInternalAvg:" in ln for ln in lines)


REF_B200 = os.path.join(ROOT, "tests", "integration", "_build", "ref_harness_b200")


@pytest.mark.parametrize("tier", ["fast", "plain"])
@pytest.mark.parametrize("case", ["lap2d_120", "stencil27_20", "skewed", "mixed_rows"])
def test_unmodified_reference_harness_runs_dropin_kernels(tmp_path, case, tier):
    """test/SpMV_test.cu of the reference, UNMODIFIED, linked with dropin/b200_SpMV_CUDA.cu and
    dropin/b200_cudaUtils.cu in place of src/SpMV_CUDA.cu and src/commons/cudaUtils.cu: it uploads with our
    spMatCpy*, launches our __global__ kernels under ITS launch geometry (incl. the stale (32,32) shape for the
    1-D ELL kernels) and checks every repetition against its own sgemvSerial with its own doubleVectorsDiff.
    Exit code 0 = every CUDA and OpenMP implementation matched.
    tier "fast": the uploaders hide this engine's narrow view (32-bit ids, SELL slices / column-major ELL, row lengths) behind the
    reference's device layout and the kernels run on it; "plain" (B200_DROPIN_PLAIN=1): the 64-bit compatibility walk."""
    if not os.path.exists(REF_B200):
        pytest.skip("reference harness not built (needs /root/reference at build time)")
    import spmv_openmp_cuda_b200 as sp
    sp.capi.require_device()
    m = {"lap2d_120": lambda: sp.synth.host_csr(sp.synth.lap2d(120)),
         "stencil27_20": lambda: sp.synth.host_csr(sp.synth.stencil27(20)),
         "skewed": lambda: sp.synth.rmat_host_csr(11, 10),  # rows longer than 256: the warp-per-long-row phase
         "mixed_rows": lambda: sp.synth.host_csr(sp.synth.mixed(20000, 40, 0.05))}[case]()
    rows = np.repeat(np.arange(m.M), np.diff(m.IRP).astype(np.int64))
    p = str(tmp_path / (case + ".mtx"))
    _write_mtx(p, m.M, m.N, rows, m.JA, m.AS)
    xv = str(tmp_path / "x.raw")
    sp.synth.host_vector(m.N).tofile(xv)
    env = dict(os.environ, OMP_SCHEDULE="nonmonotonic:static", GRID_ROWS="8", GRID_COLS="4")
    if tier == "plain":
        env["B200_DROPIN_PLAIN"] = "1"
    out = subprocess.run([REF_B200, p, xv], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    cuda_lines = [ln for ln in out.stdout.replace("\x1b[0m", "").splitlines() if ln.startswith("cudaBlockSize:")]
    assert len(cuda_lines) == 5, out.stdout[-2000:]  # 2 CSR + 3 ELL kernels (src/include/SpMV.h:130-140)


CLI = os.path.join(ROOT, "tests", "integration", "_build", "b200_main")


@pytest.mark.parametrize("case", ["lap2d_100", "skewed"])
def test_cli_driver_b200_modes(tmp_path, case):
    """tests/integration/b200_main -- the reference's command line (src/main.cu:69-139) with the engine behind it: every B200 mode
    and the reference's own CUDA_* mode strings, each checked (--check) against the reference's sgemvSerial with the strict
    comparator of the C ABI and with doubleVectorsDiff; the final stdout line keeps the reference's format (src/main.cu:268-269)."""
    if not os.path.exists(CLI):
        pytest.skip("b200_main not built (needs /root/reference at build time)")
    import spmv_openmp_cuda_b200 as sp
    sp.capi.require_device()
    m = sp.synth.host_csr(sp.synth.lap2d(100)) if case == "lap2d_100" else sp.synth.rmat_host_csr(11, 10)
    rows = np.repeat(np.arange(m.M), np.diff(m.IRP).astype(np.int64))
    p = str(tmp_path / (case + ".mtx"))
    _write_mtx(p, m.M, m.N, rows, m.JA, m.AS)
    env = dict(os.environ, OMP_SCHEDULE="nonmonotonic:static")
    modes = ["B200_CSR_ROWS", "B200_CSR_ROWS_WARP", "B200_CSR_ADAPTIVE", "B200_CSR_SELL", "B200_ELL_ROWS", "B200_ELL_ROWS_NN_TRANSPOSED",
             "B200_ELL_ROWS_WARP_NN_TRANSPOSED", "CUDA_CSR_ROWS", "CUDA_CSR_ROWS_WARP", "CUDA_ELL_ROWS", "CUDA_ELL_ROWS_WARP_NN_TRANSPOSED"]
    if m.MAX_ROW_NZ <= 255:
        modes.append("B200_CSR_XWINDOW")
    seen = {}
    for mode in modes:
        out = subprocess.run([CLI, p, "RNDVECT", mode, "--check"], capture_output=True, text=True, timeout=300, env=env)
        assert out.returncode == 0, (mode, out.stdout[-800:], out.stderr[-800:])
        last = [ln for ln in out.stdout.splitlines() if ln.startswith("cmode:")]
        assert len(last) == 1 and "elapsedInternal" in last[0], out.stdout[-500:]
        assert "strict rows failing 0" in out.stdout and "doubleVectorsDiff ok" in out.stdout, out.stdout[-500:]
        seen[mode] = int(last[0].split()[0].split(":")[1])
        assert float(last[0].split("elapsedInternal")[1]) > 0  # the kernel's CUDA-event time reached ElapsedInternal
    # prefix matching: the longer names are not swallowed by their prefixes, and CUDA_* map onto the same kinds
    assert seen["B200_CSR_ROWS"] != seen["B200_CSR_ROWS_WARP"] and seen["CUDA_CSR_ROWS"] == seen["B200_CSR_ROWS"]
    assert seen["CUDA_ELL_ROWS_WARP_NN_TRANSPOSED"] == seen["B200_ELL_ROWS_WARP_NN_TRANSPOSED"] != seen["B200_ELL_ROWS"]
