#!/usr/bin/env python
"""Does relabelling the columns by popularity (hottest columns contiguous, so that four hot x entries share a 32-byte sector and the
hot part of x fits in L1) speed the gather-bound kernels up on R-MAT?  The same matrix twice: as generated, and with JA replaced by
rank[JA] (entry order kept, i.e. the serial summation order is unchanged; the x permutation a real implementation would add is one
gather of N doubles per SpMV, timed here as well).   python tools/relabel_probe.py [scale]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import kbench  # noqa: E402
import spmv_openmp_cuda_b200 as sp  # noqa: E402
from spmv_openmp_cuda_b200 import synth  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
t0 = time.time()
mat = synth.rmat_host_csr(scale, 16)
print("# host R-MAT scale %d: M=%d nnz=%d in %.1f s" % (scale, mat.M, mat.NZ, time.time() - t0), flush=True)
ja = np.asarray(mat.JA)
cnt = np.bincount(ja.astype(np.int64), minlength=mat.N)
order = np.argsort(-cnt, kind="stable")            # columns by decreasing popularity
rank = np.empty(mat.N, dtype=np.uint64)
rank[order] = np.arange(mat.N, dtype=np.uint64)
cs = np.cumsum(cnt[order]) / float(mat.NZ)
for h in (4096, 8192, 16384, 29184, 65536, 262144):
    print("# hottest %7d columns cover %.3f of the gathers" % (h, cs[h - 1]), flush=True)
irp = np.asarray(mat.IRP).astype(np.int64)
rows = np.repeat(np.arange(mat.M, dtype=np.int64), np.diff(irp))
ja_new = rank[ja.astype(np.int64)]
t1 = time.time()
perm = np.lexsort((ja_new, rows))                  # inside every row: hottest column first (changes the summation order: tolerance kinds only)
print("# rows re-sorted by popularity rank in %.1f s" % (time.time() - t1), flush=True)
variants = [("as generated", ja, mat.AS)]
if "--skip-plain-relabel" not in sys.argv:
    variants.append(("columns relabelled by popularity", ja_new, mat.AS))
variants.append(("relabelled + hottest column first in every row", ja_new[perm], np.asarray(mat.AS)[perm]))
for label, JA, AS in variants:
    m2 = sp.Spmat.csr(mat.N, mat.IRP, np.ascontiguousarray(JA, dtype=np.uint64), AS)
    d = sp.spMatCpyCSR(m2)
    kbench.bench("rmat s%d %s" % (scale, label), d, kbench.CSR_KINDS, 20, False)
    d.free()
# the x permutation a real implementation pays per SpMV: x'[rank] = x[col]
import torch  # noqa: E402
x = torch.rand(mat.N, dtype=torch.float64, device="cuda")
idx = torch.from_numpy(order.astype(np.int64)).cuda()
for _ in range(3):
    xp = x[idx]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    xp = x[idx]
e1.record()
torch.cuda.synchronize()
print("# x permutation (torch index gather of %d doubles, 64-bit ids): %.1f us" % (mat.N, e0.elapsed_time(e1) / 20 * 1e3), flush=True)
