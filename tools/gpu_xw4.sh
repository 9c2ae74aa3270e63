#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_xw4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_xw4.log
SPMVB200_VERBOSE=1 timeout 900 python tools/kbench.py cfg1 cfg2 cfg3 cfg4s cfg5 --reps 10 > gpurun_out/kbench_xw4.log 2>&1; echo "kbench rc=$?"
grep -E "adapt" gpurun_out/kbench_xw4.log | cut -c1-150
