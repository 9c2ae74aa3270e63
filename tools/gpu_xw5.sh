#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/xw5.log; : > $L
run() { env "$@" >> $L 2>&1; }
(timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "xwin or golden or scaled or edge or delivery or iterated") > gpurun_out/pytest_xw5.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_xw5.log
for mode in 0 1; do
run SPMVB200_XW_MODE=$mode python tools/xwbench.py cfg4s 2048:8192 2048:12288 4096:8192
run SPMVB200_XW_MODE=$mode SPMVB200_XW_NO_PREFETCH=1 python tools/xwbench.py cfg4s 2048:8192
run SPMVB200_XW_MODE=$mode python tools/xwbench.py cfg4n 2048:8192
run SPMVB200_XW_MODE=$mode python tools/xwbench.py cfg2 2048:8192
done
run SPMVB200_XW_MODE=0 python tools/xwbench.py cfg4 2048:8192
run python tools/xwbench.py cfg1 2048:8192 2048:4096 --flush
grep -v "^{" $L | grep -v "^#"
