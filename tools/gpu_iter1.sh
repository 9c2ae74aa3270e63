#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "iterated or delivery or xwin") > gpurun_out/pytest_iter1.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_iter1.log
