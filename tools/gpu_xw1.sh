#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "xwin or golden or scaled or edge or adapters or pipelined") > gpurun_out/pytest_xw1.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_xw1.log
(timeout 600 python tools/kbench.py cfg4s cfg2 cfg1 --reps 10) > gpurun_out/kbench_xw1.log 2>&1; echo "kbench rc=$?"; grep -E "xwin|adapt|tiles" gpurun_out/kbench_xw1.log
