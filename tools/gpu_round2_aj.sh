#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -1
for w in cfg4s cfg4n cfg2 cfg4; do
  timeout 300 python tools/xwbench.py $w 2048:8192 --reps 20 2>&1 | grep -v "^#" | sed "s/nw=- u=- nbuf=-  *//" | tee -a $O/r02aj_xwbench.log
done
