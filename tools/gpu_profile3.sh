#!/bin/bash
mkdir -p gpurun_out
K1="python tools/kbench.py cfg1 --reps 3"
$K1 > gpurun_out/p3_plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"csr_stream|ell_colmajor" -s 2 -c 6 -f -o gpurun_out/prof_cfg1 $K1 > gpurun_out/p3_ncu1.log 2>&1
echo "ncu cfg1 rc=$?"
python tools/cfg5_sweep.py > gpurun_out/cfg5_sweep.log 2>&1; tail -3 gpurun_out/cfg5_sweep.log
