#!/bin/bash
# round 2, call 12: phase counters of the Cholesky-QR algebra kernel; stage-1 totals with it
mkdir -p gpurun_out
SVDB200_LIB=$PWD/svdsolver_b200/libsvdb200_timing.so timeout 300 python tools/panel_chol_timing.py f64 64 4096 > gpurun_out/r2_chol_timing.log 2>&1
SVDB200_LIB=$PWD/svdsolver_b200/libsvdb200_timing.so timeout 300 python tools/panel_chol_timing.py f64 32 3840 >> gpurun_out/r2_chol_timing.log 2>&1
SVDB200_LIB=$PWD/svdsolver_b200/libsvdb200_timing.so timeout 300 python tools/panel_chol_timing.py f32 64 16384 >> gpurun_out/r2_chol_timing.log 2>&1
cat gpurun_out/r2_chol_timing.log
for cfg in "3840 32 f64" "16384 64 f64" "16384 64 f32"; do
  timeout 300 python tools/stage1_only.py $cfg 2>&1 | grep "stage1 ms" | sed "s/^/chol $cfg: /"
  SVDB200_PANEL_CHOL=0 timeout 300 python tools/stage1_only.py $cfg 2>&1 | grep "stage1 ms" | sed "s/^/blk  $cfg: /"
done 2>&1 | tee gpurun_out/r2_chol_stage1.log
