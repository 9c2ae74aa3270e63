#!/bin/bash
# visit I: host pipeline without time-stamped events (schedule x chunks), x-window sensitivity to window traffic (timing-only probe)
python tools/pipe_sched_probe.py --timeline > gpurun_out/r02i_pipe_sched2.log 2>&1
python tools/xwbench.py cfg4s 2048:8192 --reps 20 > gpurun_out/r02i_xw_halfwin.log 2>&1
SPMVB200_XW_PROBE_HALFWIN=1 python tools/xwbench.py cfg4s 2048:8192 --reps 20 >> gpurun_out/r02i_xw_halfwin.log 2>&1
SPMVB200_XW_U=4 python tools/xwbench.py cfg4s 2048:8192 --reps 20 >> gpurun_out/r02i_xw_halfwin.log 2>&1
SPMVB200_XW_U=6 python tools/xwbench.py cfg4s 2048:8192 --reps 20 >> gpurun_out/r02i_xw_halfwin.log 2>&1
cat gpurun_out/r02i_pipe_sched2.log gpurun_out/r02i_xw_halfwin.log
