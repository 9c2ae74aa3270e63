#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/xw6.log; : > $L
run() { env "$@" >> $L 2>&1; }
(timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "xwin or golden or scaled or edge or delivery or iterated or unsorted or adaptive") > gpurun_out/pytest_xw6.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_xw6.log
run python tools/xwbench.py cfg4s 4096:8192 2048:8192 4096:12288
run SPMVB200_XW_U=4 python tools/xwbench.py cfg4s 4096:8192
run SPMVB200_XW_U=8 python tools/xwbench.py cfg4s 4096:8192
run SPMVB200_XW_LAYOUT=1 python tools/xwbench.py cfg4s 2048:8192
run python tools/xwbench.py cfg4n 4096:8192 2048:8192
run python tools/xwbench.py cfg2 4096:8192 2048:8192
run python tools/xwbench.py cfg1 4096:8192 2048:8192 --flush
grep -v "^{" $L | grep -v "^#"
