#!/bin/bash
# round 2, call 10: panel kernel v4 (slice in shared memory, rolled row loop): parity, single-panel timings (tall), breakdown
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py -x -q -m gpu -k "panel or tall or svdvals_chain or onestage or dist_driver" > gpurun_out/r2_t_panel4.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t_panel4.log
timeout 600 python tools/panel_only.py f32 64 4096 16384 32768 65536 > gpurun_out/r2_panel_only.log 2>&1
timeout 600 python tools/panel_only.py f64 64 4096 16384 >> gpurun_out/r2_panel_only.log 2>&1
timeout 600 python tools/panel_only.py f64 32 1024 3840 >> gpurun_out/r2_panel_only.log 2>&1
timeout 600 python tools/panel_only.py f32 32 3840 >> gpurun_out/r2_panel_only.log 2>&1
SVDB200_LIB=$PWD/svdsolver_b200/libsvdb200_timing.so timeout 600 python tools/panel_blk_timing.py 3840 32 f64 4096 64 f64 8192 64 f32 > gpurun_out/r2_blk_timing6.log 2>&1
timeout 600 python tools/panel_diag.py > gpurun_out/r2_panel_diag5.log 2>&1
tail -6 gpurun_out/r2_t_panel4.log; cat gpurun_out/r2_panel_only.log gpurun_out/r2_blk_timing6.log; grep large gpurun_out/r2_panel_diag5.log
