#!/bin/bash
# round 2, call 19: algebra kernel with four product tiles in flight per warp: parity + phase counters
mkdir -p gpurun_out
timeout 300 python tools/panel_only.py f64 32 64 3840 > gpurun_out/r2_chol_only6.log 2>&1
timeout 300 python tools/panel_only.py f64 64 128 16384 >> gpurun_out/r2_chol_only6.log 2>&1
timeout 300 python tools/panel_only.py f32 64 65536 >> gpurun_out/r2_chol_only6.log 2>&1
timeout 300 python tools/panel_only.py f64 16 512 >> gpurun_out/r2_chol_only6.log 2>&1
timeout 300 python tools/panel_only.py f32 8 1000 >> gpurun_out/r2_chol_only6.log 2>&1
cut -c1-200 gpurun_out/r2_chol_only6.log
SVDB200_LIB=$PWD/svdsolver_b200/libsvdb200_timing.so timeout 300 python tools/panel_chol_timing.py f64 64 4096 > gpurun_out/r2_chol_timing6.log 2>&1
SVDB200_LIB=$PWD/svdsolver_b200/libsvdb200_timing.so timeout 300 python tools/panel_chol_timing.py f64 32 3840 >> gpurun_out/r2_chol_timing6.log 2>&1
cat gpurun_out/r2_chol_timing6.log
