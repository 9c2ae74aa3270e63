#!/bin/bash
# writes gpurun_out/b200_harness_lap2d_150.log: the C harness (reference parser / oracle / comparator around the b200SpMV* adapters) on a
# 150 x 150 Laplacian -- the fixture tests/test_capi_cpu.py feeds to the reference's scripts/parseLog.py
mkdir -p gpurun_out
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import spmv_openmp_cuda_b200 as sp
m = sp.synth.host_csr(sp.synth.lap2d(150))
rows = np.repeat(np.arange(m.M), np.diff(m.IRP).astype(np.int64))
with open('/tmp/lap2d_150.mtx', 'w') as f:
    f.write("%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (m.M, m.N, m.NZ))
    for r, c, v in zip(rows, m.JA, m.AS):
        f.write("%d %d %.17g\n" % (r + 1, c + 1, v))
PY
R=$PWD
(cd /tmp && OMP_SCHEDULE=nonmonotonic:static $R/tests/integration/_build/b200_harness lap2d_150.mtx) > gpurun_out/b200_harness_lap2d_150.log 2>&1
tail -2 gpurun_out/b200_harness_lap2d_150.log
