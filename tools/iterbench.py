#!/usr/bin/env python
"""Iterated SpMV on one GPU (x <- A x): time per SpMV with the launch pair replayed from a CUDA graph vs plain launches vs
one-launch-at-a-time event timing.  python tools/iterbench.py [cfg1 cfg2]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import kbench  # noqa: E402
import spmv_openmp_cuda_b200 as sp  # noqa: E402
from spmv_openmp_cuda_b200 import synth  # noqa: E402


def run(label, dm, kind, iters=400):
    a, b = sp.DeviceVector(dm.N), sp.DeviceVector(dm.N)
    synth.device_vector_fill(a, dm.N)
    B = dm.algorithmic_bytes
    sp.time_kernel(kind, dm, a, b, reps=3)  # first use of a self-tuning kind tunes here
    single = float(sp.time_kernel(kind, dm, a, b, reps=25).mean()) * 1e3
    out = []
    for graph in (True, False):
        synth.device_vector_fill(a, dm.N)  # values grow with the iteration count: restart from a bounded x
        sp.iterate(kind, dm, a, b, 20, use_graph=graph)
        synth.device_vector_fill(a, dm.N)
        ms = sp.iterate(kind, dm, a, b, iters, use_graph=graph)
        out.append(ms / iters * 1e3)
    print("%-34s single-launch %7.2f us (%.3f)   iterated, graph %7.2f us (%.3f)   iterated, plain launches %7.2f us (%.3f)" % (
        label, single, B / single / 1e3 / kbench.PEAK, out[0], B / out[0] / 1e3 / kbench.PEAK, out[1], B / out[1] / 1e3 / kbench.PEAK), flush=True)


def main():
    which = sys.argv[1:] or ["cfg1", "cfg2"]
    print(sp.capi.device_info(), "peak", kbench.PEAK, "(fraction of the HBM roofline in parentheses; cfg1's 84 MB stay in L2 between iterations)")
    for w in which:
        d = synth.device_csr(synth.lap2d(1024) if w == "cfg1" else synth.stencil27(128))
        run(w + " CSR rows (stream kernel)", d, sp.CSR_ROWS)
        run(w + " CSR adaptive", d, sp.CSR_ADAPTIVE)
        e = d.to_ell(sp.FMT_ELL_COLMAJOR)
        run(w + " ELL column-major", e, sp.ELL_ROWS)
        xw = d.to_xwin()
        run(w + " x-window", xw, sp.XWIN_ROWS)


if __name__ == "__main__":
    main()
