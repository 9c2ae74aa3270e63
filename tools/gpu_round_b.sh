#!/bin/bash
# One GPU-box visit (second half of round 1): parity tests, bench lines, ncu launch list + full capture of the headline kernel, kernel benches.
TAG=${1:-r01b}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
(timeout 600 python bench.py) > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_$TAG.json
(timeout 300 python bench.py --impl reference --steps 20 --warmup 3) > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cut -c1-300 gpurun_out/bench_ref_$TAG.json
BCMD="python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 5"
$BCMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $BCMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launches rc=$?"
$BCMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ell_colmajor -s 5 -c 2 -f -o gpurun_out/prof_ell_$TAG $BCMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full ell rc=$?"
SPMVB200_VERBOSE=1 timeout 900 python tools/kbench.py cfg1 cfg2 cfg3 cfg4s cfg5 cfg4 --reps 25 > gpurun_out/kbench_$TAG.log 2>&1; echo "kbench rc=$?"
python tools/iterbench.py > gpurun_out/iterbench_$TAG.log 2>&1; echo "iterbench rc=$?"
python __graft_entry__.py --smoke 2>&1 | tail -2
