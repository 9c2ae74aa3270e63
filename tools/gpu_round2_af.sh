#!/bin/bash
# visit AF / AI: ncu --set full of the x-window kernel on the full cfg4 matrix (a steady-state launch: the first seven are the first-use
# pick's) after a clean run of the same command; launch list of the bench command
O=gpurun_out
mkdir -p $O
python tools/ncu_target.py cfg4 csr_rows > $O/r02ai_ncu_target_clean.log 2>&1; echo "clean rc=$?"; tail -1 $O/r02ai_ncu_target_clean.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:xwin_kernelILi32ELi2ELi5E -s 8 -c 1 -f -o $O/r02ai_xwin_cfg4 python tools/ncu_target.py cfg4 csr_rows > $O/r02ai_ncu.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02ai_bench_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-side --e2e-steps 3 --e2e-blocks 1 > $O/r02ai_ncu2.log 2>&1; echo "ncu list rc=$?"
