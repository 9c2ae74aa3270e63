#!/usr/bin/env python
"""BASELINE.json configs[4]: ELL vs CSR on mixed short/long rows (M = 2^23, short rows 4 nnz, a fraction p of rows
K_max long, uniform random columns), padding-ratio sweep, with and without the rowLens early exit.
Run three times (process-wide developer knobs):  python tools/cfg5_sweep.py ; SPMVB200_ELL_NO_SELL=1 python tools/cfg5_sweep.py ell
(the column-major kernel alone, without the SELL copy ELL_ROWS may build at first use) ; SPMVB200_ELL_NO_EARLY_EXIT=1 python tools/cfg5_sweep.py ell"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spmv_openmp_cuda_b200 as sp
from spmv_openmp_cuda_b200 import synth

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6551.4
only_ell = len(sys.argv) > 1 and sys.argv[1] == "ell"
noexit = os.environ.get("SPMVB200_ELL_NO_EARLY_EXIT") is not None
M = 1 << 23
print("# cfg5 sweep, M=N=2^23, short rows 4 nnz; peak %.0f GB/s; ELL early exit: %s" % (PEAK, "OFF (all K slots)" if noexit else "on"))
print("# %-5s %-6s %10s %8s | %-22s %9s %9s %7s" % ("Kmax", "p", "nnz", "pad", "kernel", "us", "GB/s alg", "frac"))
for kmax in (8, 16, 32, 48):
    for p in (0.0, 0.01, 0.05, 0.2, 0.5, 1.0):
        d = synth.device_csr(synth.mixed(M, kmax, p))
        pad = d.M * (kmax if p > 0 else 4) / d.NZ
        dx = sp.DeviceVector(d.N); synth.device_vector_fill(dx, d.N); dy = sp.DeviceVector(d.M)
        rows = []
        if not only_ell:
            sp.time_kernel(sp.CSR_ADAPTIVE, d, dx, dy, reps=2)
            t = float(sp.time_kernel(sp.CSR_ADAPTIVE, d, dx, dy, reps=10).mean())
            rows.append(("csr_adaptive[%s]" % d.adaptive_choice, t, d.algorithmic_bytes))
        e = d.to_ell(sp.FMT_ELL_COLMAJOR)
        sp.time_kernel(sp.ELL_ROWS, e, dx, dy, reps=2)
        t = float(sp.time_kernel(sp.ELL_ROWS, e, dx, dy, reps=10).mean())
        rows.append(("ell_rows[%s]" % e.exact_choice + ("(no exit)" if noexit else ""), t, e.algorithmic_bytes))
        if not only_ell:
            sl = d.to_sell()
            sp.time_kernel(sp.SELL_ROWS, sl, dx, dy, reps=2)
            t = float(sp.time_kernel(sp.SELL_ROWS, sl, dx, dy, reps=10).mean())
            rows.append(("sell32 (slots/nnz %.2f)" % ((sl.device_bytes - sl.M * 8) / 12 / max(sl.NZ, 1)), t, sl.algorithmic_bytes))
            sl.free()
        for name, t, B in rows:
            print("  %-5d %-6.2f %10d %8.2f | %-26s %9.1f %9.1f %7.3f" % (kmax, p, d.NZ, pad, name, t * 1e3, B / t / 1e6, B / t / 1e6 / PEAK), flush=True)
        d.free(); e.free()
