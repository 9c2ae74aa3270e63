#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/xw2.log; : > $L
run() { env "$@" >> $L 2>&1; }
run python tools/xwbench.py cfg4s 2048:8192 2048:12288 1024:8192 4096:8192
run SPMVB200_XW_U=3 python tools/xwbench.py cfg4s 2048:8192
run SPMVB200_XW_U=5 python tools/xwbench.py cfg4s 2048:8192
run SPMVB200_XW_U=6 python tools/xwbench.py cfg4s 1024:8192
run SPMVB200_XW_DBG=1 python tools/xwbench.py cfg4s 2048:8192
run SPMVB200_XW_DBG=2 python tools/xwbench.py cfg4s 2048:8192
run SPMVB200_XW_DBG=3 python tools/xwbench.py cfg4s 2048:8192
run SPMVB200_XW_NW=16 python tools/xwbench.py cfg4s 2048:8192 1024:8192 4096:8192
run SPMVB200_XW_NW=16 SPMVB200_XW_U=3 python tools/xwbench.py cfg4s 2048:8192
run SPMVB200_XW_NW=16 SPMVB200_XW_U=6 python tools/xwbench.py cfg4s 1024:8192
run python tools/xwbench.py cfg4n 2048:8192
run SPMVB200_XW_DBG=3 python tools/xwbench.py cfg4n 2048:8192
run python tools/xwbench.py cfg2 2048:8192
run SPMVB200_XW_DBG=3 python tools/xwbench.py cfg2 2048:8192
grep -v "^{" $L
CMD1="python tools/xwbench.py cfg4s 2048:8192 --reps 2"
$CMD1 > gpurun_out/xw3_plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:xwin_kernel -s 3 -c 1 -f -o gpurun_out/prof_xwin_c $CMD1 > gpurun_out/xw3_ncu1.log 2>&1
echo "ncu1 rc=$?"
