#!/bin/bash
mkdir -p gpurun_out
CMD1="python tools/xwbench.py cfg4s 2048:8192 --reps 2"
$CMD1 > gpurun_out/xw3_plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:xwin_kernel -s 3 -c 1 -f -o gpurun_out/prof_xwin_f $CMD1 > gpurun_out/xw3_ncu1.log 2>&1
echo "ncu1 rc=$?"
export SPMVB200_XW_NCTA=148
$CMD1 > gpurun_out/xw3_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:xwin_kernel -s 3 -c 1 -f -o gpurun_out/prof_xwin_g $CMD1 > gpurun_out/xw3_ncu2.log 2>&1
echo "ncu2 rc=$?"
