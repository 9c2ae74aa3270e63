#!/usr/bin/env python
"""Reference harness on the B200: the reference's own CUDA kernels (ref_harness_orig) next to this repo's
strict drop-ins (ref_harness_b200), same unmodified test/SpMV_test.cu, same matrix, its own host-stopwatch
timing.  Prints the per-implementation timeAvg lines of both.  Usage: python tools/ref_gpu_compare.py [n_lap] [n_stencil]"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spmv_openmp_cuda_b200 as sp  # noqa: E402

B = os.path.join(ROOT, "tests", "integration", "_build")


def write_mtx(path, m):
    rows = np.repeat(np.arange(m.M), np.diff(m.IRP).astype(np.int64)) + 1
    with open(path, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (m.M, m.N, m.NZ))
        np.savetxt(f, np.column_stack([rows, m.JA.astype(np.int64) + 1, m.AS]), fmt="%d %d %.17g")


def main():
    n_lap = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    n_st = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    env = dict(os.environ, OMP_SCHEDULE="nonmonotonic:static", GRID_ROWS="64", GRID_COLS="4")
    with tempfile.TemporaryDirectory() as d:
        for name, spec in (("lap2d_%d" % n_lap, sp.synth.lap2d(n_lap)), ("stencil27_%d" % n_st, sp.synth.stencil27(n_st))):
            m = sp.synth.host_csr(spec)
            p, xv = os.path.join(d, name + ".mtx"), os.path.join(d, "x.raw")
            write_mtx(p, m)
            sp.synth.host_vector(m.N).tofile(xv)
            for exe, tier in (("ref_harness_orig", ""), ("ref_harness_b200", "fast tier (narrow view)"), ("ref_harness_b200", "plain tier (B200_DROPIN_PLAIN=1)")):
                e2 = dict(env, B200_DROPIN_PLAIN="1") if tier.startswith("plain") else env
                out = subprocess.run([os.path.join(B, exe), p, xv], capture_output=True, text=True, env=e2)
                print("=== %s %s  %s  M=%d NZ=%d  rc=%d" % (exe, tier, name, m.M, m.NZ, out.returncode))
                lab = ""
                for ln in out.stdout.replace("\x1b[0m", "").splitlines():
                    if "@computing" in ln:
                        lab = ln.replace("\x1b[1m\x1b[92m", "").replace("\x1b[0m", "").split("func:")[1].split(" at:")[0].strip()
                    if ln.startswith("cudaBlockSize:") or ln.startswith("threadNum:"):
                        t = float(ln.split("timeAvg:")[1].split()[0])
                        print("  %-12s timeAvg %.3e s  %8.1f GFLOP/s   | %s" % (lab, t, 2 * m.NZ / t / 1e9, ln.split("\t\t")[0][:60]))
                if out.returncode:
                    print(out.stderr[-500:])


if __name__ == "__main__":
    main()
