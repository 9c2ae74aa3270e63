#!/usr/bin/env python
"""One workload, one kind, a few launches -- the program ncu wraps for a `--set full` capture of a single kernel.
    python tools/ncu_target.py cfg4 csr_rows        # x-window kernel on the full 2^30-nnz banded matrix
    python tools/ncu_target.py cfg5 csr_rows 32 0.15   # SELL kernel on mixed rows (K_max, p_long)
    python tools/ncu_target.py cfg3 csr_adapt       # R-MAT scale 22: SELL hybrid + per-row kernels
Prints the pick and the CUDA-event time of the launches (not under ncu: run it once without the profiler first)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spmv_openmp_cuda_b200 as sp  # noqa: E402
from spmv_openmp_cuda_b200 import synth  # noqa: E402

KINDS = {"csr_rows": sp.CSR_ROWS, "csr_warp": sp.CSR_ROWS_WARP, "csr_adapt": sp.CSR_ADAPTIVE, "ell_rows": sp.ELL_ROWS,
         "sell_rows": sp.SELL_ROWS, "xwin_rows": sp.XWIN_ROWS}


def main():
    cfg, kname = sys.argv[1], sys.argv[2]
    kind = KINDS[kname]
    if cfg == "cfg1":
        d = synth.device_csr(synth.lap2d(1024))
    elif cfg == "cfg2":
        d = synth.device_csr(synth.stencil27(128))
    elif cfg == "cfg3":
        d = synth.rmat_device_csr(int(sys.argv[3]) if len(sys.argv) > 3 else 22, 16)
    elif cfg == "cfg4s":  # one GPU's slice of cfg4 in an 8-GPU run: rows [0, 2^22) of the 2^25-row matrix
        d = synth.device_csr(synth.banded(1 << 25, 32, 1 << 15), 0, 1 << 22)
    elif cfg == "cfg4":
        d = synth.device_csr(synth.banded(1 << 25, 32, int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 15))
    else:
        K = int(sys.argv[3]) if len(sys.argv) > 3 else 32
        p = float(sys.argv[4]) if len(sys.argv) > 4 else 0.15
        d = synth.device_csr(synth.mixed(1 << 23, K, p))
    dm = d
    if kind == sp.ELL_ROWS:
        dm = d.to_ell(sp.FMT_ELL_COLMAJOR)
    elif kind == sp.SELL_ROWS:
        dm = d.to_sell()
    elif kind == sp.XWIN_ROWS:
        dm = d.to_xwin()
    dx, dy = sp.DeviceVector(dm.N), sp.DeviceVector(dm.M)
    synth.device_vector_fill(dx, dm.N)
    reps = int(os.environ.get("NCU_TARGET_REPS", "5"))
    sp.time_kernel(kind, dm, dx, dy, reps=2)
    t = sp.time_kernel(kind, dm, dx, dy, reps=reps, flush_l2=(cfg == "cfg1"))
    pick = dm.adaptive_choice if kind == sp.CSR_ADAPTIVE else (dm.exact_choice if kind in (sp.CSR_ROWS, sp.ELL_ROWS) else "")
    B = dm.algorithmic_bytes
    print("%s %s pick=%s M=%d NZ=%d  mean %.3f us  min %.3f us  algorithmic %.1f MB  %.1f GB/s" % (
        cfg, kname, pick, dm.M, dm.NZ, float(np.mean(t)) * 1e3, float(np.min(t)) * 1e3, B / 1e6, B / float(np.mean(t)) / 1e6))


if __name__ == "__main__":
    main()
