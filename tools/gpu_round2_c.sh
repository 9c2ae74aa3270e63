#!/bin/bash
# round-2 GPU visit C (1 GPU): L1-policy A/B on the gather-bound configs, drop-in kernels under CUDA events, cfg1 speculative batch,
# x-window recapture (steady-state launch), e2e pipeline timeline
O=gpurun_out
mkdir -p $O
for pol in 0 1; do
  for t in "cfg3 csr_adapt" "cfg3 csr_rows" "cfg3 csr_warp" "cfg3 sell_rows" "cfg5 csr_rows 32 0.15" "cfg5 csr_rows 32 0.02" "cfg5 csr_warp 32 0.15"; do
    echo -n "L1_POLICY=$pol  "; SPMVB200_L1_POLICY=$pol NCU_TARGET_REPS=20 python tools/ncu_target.py $t
  done
done 2>&1 | tee $O/r02c_l1_policy_ab.log
for spec in 0 1; do
  echo "== SPMVB200_ELL_NO_SPEC=$spec"; if [ $spec = 1 ]; then export SPMVB200_ELL_NO_SPEC=1; else unset SPMVB200_ELL_NO_SPEC; fi
  NCU_TARGET_REPS=50 python tools/ncu_target.py cfg1 ell_rows; NCU_TARGET_REPS=50 python tools/ncu_target.py cfg2 ell_rows
done 2>&1 | tee $O/r02c_ell_spec_ab.log
unset SPMVB200_ELL_NO_SPEC
B=tests/integration/_build
for w in "lap2d 1024" "stencil27 64" "stencil27 128"; do
  echo "=== reference kernels (sm_100a build of src/SpMV_CUDA.cu)"; $B/dropin_bench_orig $w
  echo "=== drop-in, fast tier"; $B/dropin_bench_b200 $w
  echo "=== drop-in, plain tier"; B200_DROPIN_PLAIN=1 $B/dropin_bench_b200 $w
done 2>&1 | tee $O/r02c_dropin_bench.log
CUDA_MODULE_LOADING=EAGER timeout 900 python tools/ref_gpu_compare.py > $O/r02c_ref_gpu_compare.log 2>&1; grep -E "===|CUDA" $O/r02c_ref_gpu_compare.log | head -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xwin_kernel -s 12 -c 1 -f -o $O/r02c_xwin_cfg4 python tools/ncu_target.py cfg4 csr_rows > $O/r02c_ncu_xwin.log 2>&1
python - <<'PY' 2>&1 | tee gpurun_out/r02c_pipe_timeline.log
import os, sys, numpy as np
sys.path.insert(0, '.')
import spmv_openmp_cuda_b200 as sp
from spmv_openmp_cuda_b200 import synth
import torch
d = synth.device_csr(synth.banded(1 << 25, 32, 1 << 15))
x = torch.empty(d.N, dtype=torch.float64).pin_memory(); y = torch.empty(d.M, dtype=torch.float64).pin_memory()
x.numpy()[:] = synth.host_vector(d.N)
for i in range(6):
    sp.spmv_host(sp.CSR_ROWS, d, x, y)
os.environ["SPMVB200_PIPE_DEBUG"] = "1"
import time
t = time.perf_counter(); sp.spmv_host(sp.CSR_ROWS, d, x, y); print("one call: %.3f ms" % ((time.perf_counter() - t) * 1e3))
PY
