#!/bin/bash
# visit AF: ncu --set full of the lean x-window kernel on the full cfg4 matrix (a steady-state launch: the first three are the first-use
# pick's) after a clean run of the same command
O=gpurun_out
mkdir -p $O
python tools/ncu_target.py cfg4 csr_rows > $O/r02af_ncu_target_clean.log 2>&1; echo "clean rc=$?"; tail -1 $O/r02af_ncu_target_clean.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:xwin_kernelILi32ELi2ELi5ELb1E -s 5 -c 1 -f -o $O/r02af_xwin_cfg4_lean python tools/ncu_target.py cfg4 csr_rows > $O/r02af_ncu.log 2>&1; echo "ncu rc=$?"
