#!/bin/bash
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_exact.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_exact.log
SPMVB200_VERBOSE=1 python tools/kbench.py cfg1 cfg2 cfg3 cfg4s cfg5 --reps 10 2>&1 | grep -E "csr_rows|exact-kind" | cut -c1-200
