#!/bin/bash
# visit T: why is the x-window kernel 27 % faster per non-zero on a narrow band (3 tiles per row block) than on cfg4's (9 tiles)?
# ncu --set full of one steady-state launch each (after a clean run of the same command)
O=gpurun_out
for w in cfg4s cfg4n; do
  python tools/xwbench.py $w 2048:8192 --reps 5 > $O/r02t_xw_$w.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:xwin_kernel -s 4 -c 1 -f -o $O/r02t_xw_$w python tools/xwbench.py $w 2048:8192 --reps 5 > /dev/null 2>&1
  grep -v "^#" $O/r02t_xw_$w.log
done
ls -la $O/r02t_xw_*.ncu-rep
