#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list + full capture of the dominant kernels.
# Usage (under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
(timeout 600 python bench.py) > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json
(timeout 300 python bench.py --impl reference --steps 20 --warmup 3) > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cat gpurun_out/bench_ref_$TAG.json
BCMD="python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 5"
$BCMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $BCMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launches rc=$?"
$BCMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ell_colmajor -s 5 -c 2 -f -o gpurun_out/prof_ell_$TAG $BCMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full ell rc=$?"
KCMD="python tools/kbench.py cfg2 cfg1 --reps 3"
$KCMD > gpurun_out/plain3_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:csr_stream -s 3 -c 2 -f -o gpurun_out/prof_csr_$TAG $KCMD > gpurun_out/ncu_full_csr_$TAG.log 2>&1
echo "ncu full csr rc=$?"
ls -la gpurun_out | tail -20
