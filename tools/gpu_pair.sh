#!/bin/bash
for p in 0 2 3 4; do
echo "pair $p"; SPMVB200_ELL_PAIR=$p python tools/kbench.py cfg2 --reps 25 2>&1 | grep "ELL     ell_rows" | cut -c1-150
done
SPMVB200_ELL_PAIR=3 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or scaled or edge or full_size or delivery" 2>&1 | tail -2
python __graft_entry__.py --smoke 2>&1 | tail -1
