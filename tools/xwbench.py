#!/usr/bin/env python
"""x-window kernel sweep on one workload: python tools/xwbench.py <cfg4s|cfg4|cfg2|cfg1|cfg4n> R:W [R:W ...] [--reps N] [--flush]
Kernel shape knobs come from the environment (SPMVB200_XW_NW, SPMVB200_XW_U, SPMVB200_XW_NBUF)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import kbench  # noqa: E402
import spmv_openmp_cuda_b200 as sp  # noqa: E402
from spmv_openmp_cuda_b200 import synth  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 10
    if "--reps" in sys.argv:
        args.remove(str(reps))
    w = args[0]
    geoms = [tuple(int(v) for v in a.split(":")) for a in args[1:]]
    spec = {"cfg4s": lambda: synth.banded(1 << 22, 32, 1 << 15), "cfg4n": lambda: synth.banded(1 << 22, 32, 1 << 12),
            "cfg4": lambda: synth.banded(1 << 25, 32, 1 << 15), "cfg2": lambda: synth.stencil27(128),
            "cfg1": lambda: synth.lap2d(1024)}[w]()
    d = synth.device_csr(spec)
    tag = "%s nw=%s u=%s nbuf=%s" % (w, os.environ.get("SPMVB200_XW_NW", "-"), os.environ.get("SPMVB200_XW_U", "-"), os.environ.get("SPMVB200_XW_NBUF", "-"))
    kbench.bench_xwin(tag, d, geoms, reps, "--flush" in sys.argv)


if __name__ == "__main__":
    main()
