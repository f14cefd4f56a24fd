#!/bin/bash
# round 2, call 16: algebra kernel v6/v7 (three chains per step: Cholesky, sqrt helper, LU two steps behind; unconditional updates)
mkdir -p gpurun_out
timeout 300 python tools/panel_only.py f64 32 64 3840 > gpurun_out/r2_chol_only5.log 2>&1
timeout 300 python tools/panel_only.py f64 64 128 16384 >> gpurun_out/r2_chol_only5.log 2>&1
timeout 300 python tools/panel_only.py f32 64 4096 65536 >> gpurun_out/r2_chol_only5.log 2>&1
timeout 300 python tools/panel_only.py f64 16 512 >> gpurun_out/r2_chol_only5.log 2>&1
timeout 300 python tools/panel_only.py f32 8 1000 >> gpurun_out/r2_chol_only5.log 2>&1
cut -c1-200 gpurun_out/r2_chol_only5.log
SVDB200_LIB=$PWD/svdsolver_b200/libsvdb200_timing.so timeout 300 python tools/panel_chol_timing.py f64 64 4096 > gpurun_out/r2_chol_timing5.log 2>&1
SVDB200_LIB=$PWD/svdsolver_b200/libsvdb200_timing.so timeout 300 python tools/panel_chol_timing.py f64 32 3840 >> gpurun_out/r2_chol_timing5.log 2>&1
cat gpurun_out/r2_chol_timing5.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "panel or tall or svdvals_chain or onestage or stage1" > gpurun_out/r2_t_chol5.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t_chol5.log
tail -5 gpurun_out/r2_t_chol5.log
