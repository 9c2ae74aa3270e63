#!/bin/bash
# round-2 GPU visit F (1 GPU): the hot-x hybrid -- parity test, then R-MAT scale 22 with the candidate timed next to the others
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hotx or sell_hybrid" 2>&1 | tail -5
SPMVB200_VERBOSE=1 NCU_TARGET_REPS=20 python tools/ncu_target.py cfg3 csr_adapt 2>&1 | tee $O/r02f_hotx_cfg3.log | tail -4
SPMVB200_VERBOSE=1 NCU_TARGET_REPS=20 python tools/ncu_target.py cfg3 csr_adapt 20 2>&1 | tee -a $O/r02f_hotx_cfg3.log | tail -3
SPMVB200_FORCE_CAND=14 timeout 600 ncu --set full --clock-control none --import-source on -k regex:hotx_kernel -s 3 -c 1 -f -o $O/r02f_hotx_cfg3 python tools/ncu_target.py cfg3 csr_adapt > /dev/null 2>&1
ls -la $O/r02f_hotx_cfg3.ncu-rep
