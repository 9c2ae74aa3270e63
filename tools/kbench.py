#!/usr/bin/env python
"""Developer kernel benchmark: every kernel kind on the BASELINE.json workloads, CUDA-event timed.
Usage: python tools/kbench.py [cfg1 cfg2 cfg3 cfg4 cfg5 ...] [--reps N]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spmv_openmp_cuda_b200 as sp  # noqa: E402
from spmv_openmp_cuda_b200 import synth  # noqa: E402

PEAK = 6551.4
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass

CSR_KINDS = [("csr_rows", sp.CSR_ROWS), ("csr_warp", sp.CSR_ROWS_WARP), ("csr_adapt", sp.CSR_ADAPTIVE)]


def bench(label, dm, kinds, reps, flush):
    dx = sp.DeviceVector(dm.N)
    synth.device_vector_fill(dx, dm.N)
    dy = sp.DeviceVector(dm.M)
    B = dm.algorithmic_bytes
    for name, kind in kinds:
        if not dm.supports(kind):
            continue
        sp.time_kernel(kind, dm, dx, dy, reps=3, flush_l2=False)
        t = sp.time_kernel(kind, dm, dx, dy, reps=reps, flush_l2=flush)
        tmin, tmean = float(t.min()), float(t.mean())
        if kind == sp.CSR_ADAPTIVE:
            name = name + "[" + dm.adaptive_choice + "]"
        elif kind in (sp.CSR_ROWS, sp.ELL_ROWS):
            name = name + "[" + dm.exact_choice + "]"
        print("%-28s %-24s flush=%d  min %8.3f us  mean %8.3f us  %8.1f GB/s (mean)  frac %.3f  %7.1f GFLOP/s   [M=%d NZ=%d bytes=%.1f MB]" % (
            label, name, flush, tmin * 1e3, tmean * 1e3, B / tmean / 1e6, B / tmean / 1e6 / PEAK, 2 * dm.NZ / tmean / 1e6, dm.M, dm.NZ, B / 1e6), flush=True)


def bench_xwin(label, d, geoms, reps, flush=False):
    """x-window copies of a CSR handle at several (rows per block, window columns) geometries."""
    for R, W in geoms:
        try:
            xw = d.to_xwin(R, W)
        except sp.SpmvB200Error as e:
            print("%-28s xwin R=%d W=%d: %s" % (label, R, W, e), flush=True)
            continue
        info = xw.xwin_info
        bench(label, xw, [("xwin R=%d W=%d x%d" % (R, W, info["ring"]), sp.XWIN_ROWS)], reps, flush)
        print("#   tiles=%d (%.2f per row block), moved %.1f MB of which x windows %.1f MB" % (
            info["ntiles"], info["ntiles"] / ((d.M + R - 1) // R), info["moved_bytes"] / 1e6, info["ntiles"] * W * 8 / 1e6), flush=True)
        xw.free()


XW_GEOMS = [(4096, 8192), (2048, 8192), (8192, 8192), (4096, 4096), (4096, 12288), (2048, 4096), (1024, 4096)]


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    reps = 25
    if "--reps" in sys.argv:
        reps = int(sys.argv[sys.argv.index("--reps") + 1])
    which = args or ["cfg1", "cfg2", "cfg3", "cfg5", "cfg4"]
    print(sp.capi.device_info(), "peak", PEAK)
    for w in which:
        t0 = time.time()
        if w == "cfg1":
            d = synth.device_csr(synth.lap2d(1024))
            bench("cfg1 lap2d 1024^2", d, CSR_KINDS, reps, 1)   # L2 flushed by reading (clean lines)
            bench("cfg1 lap2d 1024^2", d, CSR_KINDS, reps, 2)   # L2 flushed by a memset (dirty lines: write-back inside the timed kernel)
            bench("cfg1 lap2d 1024^2", d, CSR_KINDS, reps, 0)
            bench_xwin("cfg1 lap2d 1024^2", d, [(1024, 4096), (1024, 2048), (2048, 4096), (512, 2048)], reps, True)
            bench_xwin("cfg1 lap2d 1024^2", d, [(1024, 4096), (1024, 2048), (2048, 4096), (512, 2048)], reps, False)
            e = d.to_ell(sp.FMT_ELL_COLMAJOR)
            bench("cfg1 lap2d ELL", e, [("ell_rows", sp.ELL_ROWS)], reps, True)
        elif w == "cfg2":
            d = synth.device_csr(synth.stencil27(128))
            bench("cfg2 stencil27 128^3 CSR", d, CSR_KINDS, reps, False)
            bench_xwin("cfg2 stencil27 128^3", d, XW_GEOMS, reps)
            e = d.to_ell(sp.FMT_ELL_COLMAJOR)
            bench("cfg2 stencil27 128^3 ELL", e, [("ell_rows", sp.ELL_ROWS)], reps, False)
            e.free()
            e = d.to_ell(sp.FMT_ELL_ROWMAJOR)
            bench("cfg2 stencil27 128^3 ELLrm", e, [("ell_nt", sp.ELL_ROWS_NT), ("ell_warp_nt", sp.ELL_ROWS_WARP_NT)], reps, False)
        elif w == "cfg3":
            d = synth.rmat_device_csr(22, 16)
            bench("cfg3 rmat s22 ef16", d, CSR_KINDS, reps, False)
            for sigma in (1024, 16384, 1 << 18):
                e = d.to_sell(sigma)
                bench("cfg3 rmat s22 ef16 SELL s=%d" % sigma, e, [("sell_rows", sp.SELL_ROWS)], reps, False)
                print("#   SELL slots / nnz = %.3f" % ((e.device_bytes - e.M * 8) / 12 / max(d.NZ, 1)), flush=True)
                e.free()
        elif w == "cfg4":
            for hw in (1 << 15, 1 << 12):
                d = synth.device_csr(synth.banded(1 << 25, 32, hw))
                bench("cfg4 banded 2^25 w=%d" % hw, d, CSR_KINDS, max(5, reps // 5), False)
                bench_xwin("cfg4 banded 2^25 w=%d" % hw, d, XW_GEOMS[:3], max(5, reps // 5))
                d.free()
        elif w == "cfg4s":
            for hw in (1 << 15, 1 << 12):
                d = synth.device_csr(synth.banded(1 << 22, 32, hw))
                bench("cfg4s banded 2^22 w=%d" % hw, d, CSR_KINDS, reps, False)
                bench_xwin("cfg4s banded 2^22 w=%d" % hw, d, XW_GEOMS, reps)
                d.free()
        elif w == "cfg5":
            for kmax, p in ((32, 0.0), (32, 0.02), (32, 0.15), (32, 1.0)):
                d = synth.device_csr(synth.mixed(1 << 23, kmax, p))
                pad = d.M * kmax / max(d.NZ, 1)
                bench("cfg5 mixed K=%d p=%.2f pad=%.1f" % (kmax, p, pad), d, CSR_KINDS, reps, False)
                e = d.to_ell(sp.FMT_ELL_COLMAJOR)
                bench("cfg5 mixed K=%d p=%.2f ELL" % (kmax, p), e, [("ell_rows", sp.ELL_ROWS)], reps, False)
                e.free()
                e = d.to_sell()
                bench("cfg5 mixed K=%d p=%.2f SELL" % (kmax, p), e, [("sell_rows", sp.SELL_ROWS)], reps, False)
                d.free(); e.free()
        print("# %s done in %.1f s" % (w, time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
