#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_ell16.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ell16.log
python tools/kbench.py cfg2 cfg1 --reps 25 2>&1 | grep -E "ell_rows|ELL"
SPMVB200_ELL_NO_IDX16=1 python tools/kbench.py cfg2 cfg1 --reps 25 2>&1 | grep -E "ell_rows|ELL"
(timeout 600 python bench.py --steps 200 --warmup 10) > gpurun_out/bench_ell16.json 2> gpurun_out/bench_ell16.err; echo "bench rc=$?"; cat gpurun_out/bench_ell16.json
