for cfg in "R=2048 W=8192" "R=1024 W=8192" "R=1024 W=4096" "R=2048 W=4096" "R=4096 W=8192"; do
  eval $cfg
  echo -n "XW_R=$R XW_W=$W  "; SPMVB200_XW_R=$R SPMVB200_XW_W=$W NCU_TARGET_REPS=30 python tools/ncu_target.py cfg4s xwin_rows
done 2>&1 | tee gpurun_out/r02e_xwin_slice_geometry.log
for cfg in "R=2048 W=8192" "R=1024 W=8192"; do
  eval $cfg
  echo -n "XW_R=$R XW_W=$W  "; SPMVB200_XW_R=$R SPMVB200_XW_W=$W NCU_TARGET_REPS=30 python tools/ncu_target.py cfg2 xwin_rows
done 2>&1 | tee -a gpurun_out/r02e_xwin_slice_geometry.log
