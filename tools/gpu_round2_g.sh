#!/bin/bash
# cfg1 isolated-launch variants of the column-major ELL kernel (L2 flushed between launches)
O=gpurun_out; mkdir -p $O
for v in "PAIR=3" "PAIR=5" "PAIR=6" "PAIR=4" "PAIR=2" "PAIR=0 UNROLL=5" "PAIR=0 UNROLL=3" "PAIR=0 UNROLL=6"; do
  env=""; for kv in $v; do env="$env SPMVB200_ELL_$kv"; done
  echo -n "$v  "; env $env NCU_TARGET_REPS=100 python tools/ncu_target.py cfg1 ell_rows
done 2>&1 | tee $O/r02g_cfg1_ell_variants.log
