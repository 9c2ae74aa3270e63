#!/bin/bash
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_warp.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_warp.log
python tools/kbench.py cfg1 cfg2 cfg3 cfg4s cfg5 --reps 10 2>&1 | grep -E "csr_warp" | cut -c1-150
python tools/cfg5_sweep.py > gpurun_out/cfg5_sweep_r01c.log 2>&1; tail -30 gpurun_out/cfg5_sweep_r01c.log | cut -c1-200
