#!/bin/bash
# round 2, call 11: Cholesky-QR panel (stage1_panel_chol.cu): single-panel parity + timing, then the stage-1 parity tests with it as default
mkdir -p gpurun_out
timeout 300 python tools/panel_only.py f64 32 64 96 1024 3840 > gpurun_out/r2_chol_only.log 2>&1
timeout 300 python tools/panel_only.py f32 32 3840 >> gpurun_out/r2_chol_only.log 2>&1
timeout 300 python tools/panel_only.py f64 64 128 4096 16384 >> gpurun_out/r2_chol_only.log 2>&1
timeout 300 python tools/panel_only.py f32 64 4096 16384 65536 >> gpurun_out/r2_chol_only.log 2>&1
timeout 300 python tools/panel_only.py f64 16 512 >> gpurun_out/r2_chol_only.log 2>&1
timeout 300 python tools/panel_only.py f32 8 1000 >> gpurun_out/r2_chol_only.log 2>&1
cat gpurun_out/r2_chol_only.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "panel or tall or svdvals_chain or onestage or stage1" > gpurun_out/r2_t_chol1.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t_chol1.log
tail -15 gpurun_out/r2_t_chol1.log
