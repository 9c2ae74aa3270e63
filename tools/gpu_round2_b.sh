#!/bin/bash
# round-2 GPU visit B: tests, drop-in comparison, e2e chunk sweep, ncu captures (each command first runs without ncu)
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/r02b_pytest.log 2>&1; tail -3 $O/r02b_pytest.log
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import spmv_openmp_cuda_b200 as sp
m = sp.synth.host_csr(sp.synth.lap2d(150))
rows = np.repeat(np.arange(m.M), np.diff(m.IRP).astype(np.int64))
with open('/tmp/lap2d_150.mtx', 'w') as f:
    f.write("%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (m.M, m.N, m.NZ))
    for r, c, v in zip(rows, m.JA, m.AS):
        f.write("%d %d %.17g\n" % (r + 1, c + 1, v))
PY
(cd /tmp && OMP_SCHEDULE=nonmonotonic:static $OLDPWD/tests/integration/_build/b200_harness lap2d_150.mtx) > $O/r02b_b200_harness_lap2d_150.log 2>&1; tail -2 $O/r02b_b200_harness_lap2d_150.log
timeout 600 python tools/ref_gpu_compare.py > $O/r02b_ref_gpu_compare.log 2>&1; grep -E "===|CUDA" $O/r02b_ref_gpu_compare.log | head -40
for c in 8 16 24 32; do
  SPMVB200_HOST_CHUNKS=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-side --e2e-steps 20 --e2e-blocks 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunks $c', d['e2e']['ms_per_step'], d['e2e']['pinned']['ms_per_step'], d['e2e']['link_ceiling']['duplex_ms'])" | tee -a $O/r02b_e2e_chunks.log
done
# ncu: launch list of the bench command
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02b_bench_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-side --e2e-steps 3 --e2e-blocks 1 > $O/r02b_ncu_bench.log 2>&1
# full captures
python tools/ncu_target.py cfg4 csr_rows > $O/r02b_ncu_targets.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xwin_kernel -s 4 -c 1 -f -o $O/r02b_xwin_cfg4 python tools/ncu_target.py cfg4 csr_rows >> $O/r02b_ncu_targets.log 2>&1
python tools/ncu_target.py cfg5 csr_rows 32 0.15 >> $O/r02b_ncu_targets.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sell_kernel -s 4 -c 1 -f -o $O/r02b_sell_cfg5 python tools/ncu_target.py cfg5 csr_rows 32 0.15 >> $O/r02b_ncu_targets.log 2>&1
python tools/ncu_target.py cfg3 csr_adapt >> $O/r02b_ncu_targets.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sell_kernel|csr_midrow|csr_longrow|csr_vector" -s 12 -c 4 -f -o $O/r02b_adapt_cfg3 python tools/ncu_target.py cfg3 csr_adapt >> $O/r02b_ncu_targets.log 2>&1
python tools/ncu_target.py cfg3 csr_warp >> $O/r02b_ncu_targets.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"csr_vector" -s 2 -c 1 -f -o $O/r02b_vector_cfg3 python tools/ncu_target.py cfg3 csr_warp >> $O/r02b_ncu_targets.log 2>&1
cat $O/r02b_ncu_targets.log | grep -E "^cfg" 
ls -la $O/*.ncu-rep
