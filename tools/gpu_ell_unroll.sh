#!/bin/bash
for u in 4 3 5 6 8 9; do
echo "unroll $u"; SPMVB200_ELL_UNROLL=$u python tools/kbench.py cfg2 --reps 25 2>&1 | grep "ELL     ell_rows" | cut -c1-150
done
for u in 4 6 9; do
echo "bench loop unroll $u"; SPMVB200_ELL_UNROLL=$u python bench.py --steps 300 --warmup 10 --no-cpu --e2e-steps 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
done
