#!/bin/bash
# round 2, call 25: list-pipeline test after the like-with-like fix, Cholesky-QR panel tests after the shuffle change, phase counters
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py -q -m gpu -k "bidiagonalize_many or chol or tall or svdvals_chain" > gpurun_out/r2_t_sub7.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t_sub7.log
tail -5 gpurun_out/r2_t_sub7.log
SVDB200_LIB=$PWD/svdsolver_b200/libsvdb200_timing.so timeout 300 python tools/panel_chol_timing.py f64 64 4096 > gpurun_out/r2_chol_timing7.log 2>&1
SVDB200_LIB=$PWD/svdsolver_b200/libsvdb200_timing.so timeout 300 python tools/panel_chol_timing.py f64 32 3840 >> gpurun_out/r2_chol_timing7.log 2>&1
grep "algebra kernel\|elimination" gpurun_out/r2_chol_timing7.log
