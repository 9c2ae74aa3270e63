#!/bin/bash
(timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "adaptive or unsorted or golden or scaled or edge") > gpurun_out/pytest_sellh.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_sellh.log
SPMVB200_VERBOSE=1 python tools/kbench.py cfg3 --reps 10 2>&1 | grep -E "adapt|tuning" | cut -c1-420
python tools/kbench.py cfg2 --reps 25 2>&1 | grep -E "ELL     ell_rows" | cut -c1-150
