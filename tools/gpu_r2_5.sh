#!/bin/bash
# round 2, call 5: per-phase counters of the fast stage-2 kernel; blocked panel kernel v2 (parity, timing)
mkdir -p gpurun_out
export SVDB200_LIB=$PWD/svdsolver_b200/libsvdb200_timing.so
python tools/stage2_only.py 3840 32 f64 > gpurun_out/r2_s2f_phases.log 2>&1
python tools/stage2_only.py 3840 32 f32 >> gpurun_out/r2_s2f_phases.log 2>&1
timeout 600 python tools/panel_blk_timing.py 1920 32 f64 3840 32 f64 3840 32 f32 4096 64 f64 8192 64 f32 > gpurun_out/r2_blk_timing2.log 2>&1
unset SVDB200_LIB
timeout 900 python tools/panel_diag.py > gpurun_out/r2_panel_diag3.log 2>&1
cat gpurun_out/r2_s2f_phases.log gpurun_out/r2_blk_timing2.log gpurun_out/r2_panel_diag3.log
