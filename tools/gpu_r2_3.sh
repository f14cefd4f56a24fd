#!/bin/bash
# round 2, call 3: where does the blocked panel kernel spend its time; sign-insensitive parity numbers
mkdir -p gpurun_out
export SVDB200_LIB=$PWD/svdsolver_b200/libsvdb200_timing.so
timeout 600 python tools/panel_blk_timing.py 1920 32 f64 3840 32 f64 3840 32 f32 4096 64 f64 8192 64 f32 > gpurun_out/r2_blk_timing.log 2>&1
unset SVDB200_LIB
timeout 600 python -c "
import sys; sys.argv=['x']; sys.path.insert(0,'tools')
import panel_diag as P
P.small()
" > gpurun_out/r2_panel_diag2.log 2>&1
cat gpurun_out/r2_blk_timing.log; cat gpurun_out/r2_panel_diag2.log
