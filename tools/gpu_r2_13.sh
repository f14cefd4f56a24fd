#!/bin/bash
# round 2, call 13: algebra kernel v2 (DMMA products, owner-computed step scalars): parity, phase counters, stage-1 totals
mkdir -p gpurun_out
timeout 300 python tools/panel_only.py f64 32 64 1024 3840 > gpurun_out/r2_chol_only2.log 2>&1
timeout 300 python tools/panel_only.py f64 64 128 16384 >> gpurun_out/r2_chol_only2.log 2>&1
timeout 300 python tools/panel_only.py f32 64 4096 65536 >> gpurun_out/r2_chol_only2.log 2>&1
timeout 300 python tools/panel_only.py f64 16 512 >> gpurun_out/r2_chol_only2.log 2>&1
timeout 300 python tools/panel_only.py f32 8 1000 >> gpurun_out/r2_chol_only2.log 2>&1
cut -c1-330 gpurun_out/r2_chol_only2.log
SVDB200_LIB=$PWD/svdsolver_b200/libsvdb200_timing.so timeout 300 python tools/panel_chol_timing.py f64 64 4096 > gpurun_out/r2_chol_timing2.log 2>&1
SVDB200_LIB=$PWD/svdsolver_b200/libsvdb200_timing.so timeout 300 python tools/panel_chol_timing.py f64 32 3840 >> gpurun_out/r2_chol_timing2.log 2>&1
cat gpurun_out/r2_chol_timing2.log
for cfg in "3840 32 f64" "16384 64 f64" "16384 64 f32"; do
  timeout 300 python tools/stage1_only.py $cfg 2>&1 | grep "stage1 ms" | sed "s/^/chol $cfg: /"
done 2>&1 | tee gpurun_out/r2_chol_stage1b.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "panel or tall or svdvals_chain or onestage or stage1" > gpurun_out/r2_t_chol2.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t_chol2.log
tail -5 gpurun_out/r2_t_chol2.log
