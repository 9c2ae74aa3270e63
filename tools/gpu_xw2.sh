#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/xw2.log; : > $L
run() { env "$@" >> $L 2>&1; }
(timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "xwin or golden or scaled or edge") > gpurun_out/pytest_xw2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_xw2.log
run python tools/xwbench.py cfg4s 2048:8192 2048:12288 1024:8192 4096:8192
run SPMVB200_XW_U=4 python tools/xwbench.py cfg4s 2048:8192
run SPMVB200_XW_U=6 python tools/xwbench.py cfg4s 2048:8192
for n in 148 296 592 1024; do
run SPMVB200_XW_NCTA=$n python tools/xwbench.py cfg4s 2048:8192
done
run SPMVB200_XW_NW=16 python tools/xwbench.py cfg4s 2048:8192
run SPMVB200_XW_NW=16 SPMVB200_XW_NCTA=148 python tools/xwbench.py cfg4s 2048:8192
run SPMVB200_XW_NW=16 SPMVB200_XW_U=4 python tools/xwbench.py cfg4s 2048:8192
run SPMVB200_XW_NW=16 SPMVB200_XW_U=4 SPMVB200_XW_NCTA=148 python tools/xwbench.py cfg4s 2048:8192
run python tools/xwbench.py cfg4n 2048:8192
run SPMVB200_XW_NCTA=148 python tools/xwbench.py cfg4n 2048:8192
run python tools/xwbench.py cfg2 2048:8192 1024:8192
run SPMVB200_XW_NCTA=148 python tools/xwbench.py cfg2 2048:8192 1024:8192
run python tools/xwbench.py cfg1 1024:4096 2048:4096 --flush
run SPMVB200_XW_NCTA=148 python tools/xwbench.py cfg1 1024:4096 2048:4096 --flush
grep -v "^{" $L
