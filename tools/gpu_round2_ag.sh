#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1
for w in cfg4n cfg2; do
  timeout 300 python tools/xwbench.py $w 2048:8192 --reps 20 2>&1 | grep -v "^#" | sed "s/^/EIDX general: /;s/nw=- u=- nbuf=-  *//" | tee -a $O/r02ag_xw_eidx.log
done
SPMVB200_XW_LEAN=0 timeout 300 python tools/xwbench.py cfg4s 2048:8192 --reps 20 2>&1 | grep -v "^#" | sed "s/^/EIDX general (lean off): /;s/nw=- u=- nbuf=-  *//" | tee -a $O/r02ag_xw_eidx.log
SPMVB200_XW_LEAN=0 SPMVB200_XW_U=6 timeout 300 python tools/xwbench.py cfg4s 2048:8192 --reps 20 2>&1 | grep -v "^#" | sed "s/^/EIDX general U=6 (lean off): /;s/nw=- u=6 nbuf=-  *//" | tee -a $O/r02ag_xw_eidx.log
timeout 300 python tools/xwbench.py cfg1 1024:4096 2048:4096 --reps 20 --flush 2>&1 | grep -v "^#" | sed "s/^/EIDX general: /;s/nw=- u=- nbuf=-  *//" | tee -a $O/r02ag_xw_eidx.log
timeout 300 python tools/xwbench.py cfg2 4096:8192 1024:8192 --reps 20 2>&1 | grep -v "^#" | sed "s/^/EIDX general: /;s/nw=- u=- nbuf=-  *//" | tee -a $O/r02ag_xw_eidx.log
