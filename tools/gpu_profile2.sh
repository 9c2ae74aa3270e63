#!/bin/bash
# ncu --set full captures of the CSR kernels (after the same command ran clean without ncu).
mkdir -p gpurun_out
K1="python tools/kbench.py cfg2 --reps 3"
$K1 > gpurun_out/p2_plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:csr_stream -s 4 -c 1 -f -o gpurun_out/prof_csr_stream_rowwalk $K1 > gpurun_out/p2_ncu1.log 2>&1
echo "ncu csr_stream rc=$?"
K2="python tools/kbench.py cfg4s --reps 3"
SPMVB200_FORCE_CAND=9 $K2 > gpurun_out/p2_plain2.log 2>&1 &&
SPMVB200_FORCE_CAND=9 ncu --set full --clock-control none --import-source on -k regex:csr_vector_span -s 8 -c 1 -f -o gpurun_out/prof_csr_vspan_w4096 $K2 > gpurun_out/p2_ncu2.log 2>&1
echo "ncu vspan rc=$?"
SPMVB200_FORCE_CAND=4 ncu --set full --clock-control none --import-source on -k regex:csr_vector_kernel -s 2 -c 1 -f -o gpurun_out/prof_csr_vector_w32768 $K2 > gpurun_out/p2_ncu3.log 2>&1
echo "ncu vector rc=$?"
ls -la gpurun_out/*.ncu-rep
