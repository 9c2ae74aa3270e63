#!/bin/bash
# visit AL: every kernel kind on cfg1, cfg2, cfg3, cfg5 at the final state of round 2 (cfg4: r02aj / r02ak)
O=gpurun_out
mkdir -p $O
timeout 1200 python tools/kbench.py cfg1 cfg2 cfg3 cfg5 --reps 10 > $O/r02al_kbench_all.log 2>&1; echo "kbench rc=$?"
grep -c "frac" $O/r02al_kbench_all.log; grep "frac" $O/r02al_kbench_all.log | grep -v xwin | cut -c1-150 | head -60
