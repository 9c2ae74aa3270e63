#!/bin/bash
# ncu --set full of the x-window kernel inside the cfg4 bench line (full size, 1 GPU)
mkdir -p gpurun_out
BCMD="python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu --e2e-steps 2"
$BCMD > gpurun_out/plain_cfg4.log 2> gpurun_out/plain_cfg4.err &&
ncu --set full --clock-control none --import-source on -k regex:xwin_kernel -s 6 -c 1 -f -o gpurun_out/prof_xwin_cfg4 $BCMD > gpurun_out/ncu_cfg4.log 2>&1
echo "ncu rc=$?"
cut -c1-300 gpurun_out/plain_cfg4.log
